/*
 * wh_b200.h — C ABI of the B200-native batched warehouse hot path (libwh_b200.so).
 *
 * This is the drop-in boundary: what a maintainer of ffahleraz/rllib-warehouse would bind
 * (ctypes) in place of the numpy bodies of
 *     Warehouse.reset   warehouse/core.py:167-260      -> wh_reset (+ wh_build_obs, flavour 1)
 *     Warehouse.step    warehouse/core.py:262-442      -> wh_step  (fused with the obs build)
 *     observation dict  warehouse/core.py:224-260,371-432 -> wh_build_obs
 *     WarehouseRandomGreedySolver.compute_action  baseline/solvers.py:27-58 -> wh_greedy
 *     variant constants warehouse/variants.py:19-98    -> wh_config
 * for N independent environments at once.
 *
 * Conventions
 *  - extern "C", plain pointers and sizes only. No torch / C++ types cross this boundary.
 *  - Layer 1 (wh_reset / wh_step / wh_build_obs / wh_greedy / wh_greedy_step): every data pointer is a
 *    DEVICE pointer owned by the caller (e.g. a torch CUDA tensor); the functions launch sm_100a
 *    kernels on `stream` (a cudaStream_t passed as void*, NULL = default stream), never allocate,
 *    never synchronise, and return 0 or a non-zero error code (cudaError_t value, or WH_E_* below).
 *    There is NO CPU fallback: without a CUDA device every call fails.
 *  - Layer 2 (wh_env_*): a handle that owns device state plus pinned staging and takes HOST
 *    buffers; it pipelines H2D actions -> kernels -> D2H rewards/dones in env chunks.
 *  - Env-major layout: every tensor is [N, ...] with env e's record contiguous.
 *  - Agent rows are padded to R = num_requests; rows >= num_agents[e] hold -1 in the state and
 *    never act. Actions: 0..8 (core.py:38 MOVES), -1 = agent absent from the action dict.
 */
#ifndef WH_B200_H
#define WH_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WH_MAX_RACKS 8
#define WH_NUM_STATS 80

/* error codes outside the cudaError_t range */
#define WH_E_CONFIG 10001     /* unsupported configuration (limits below) */
#define WH_E_ARG    10002     /* NULL / inconsistent arguments */
#define WH_E_NCCL   10003     /* ncclAllReduce not resolvable in this process / NCCL error (code - 11000) */

/* Limits: R <= 32, P = 4*L*L <= 64 (L <= 4), D = 4*(dim-4) <= 64 (dim <= 20), dim <= 127.
 * All six reference variants (variants.py:19-98) are inside them. */
typedef struct wh_config {
    int32_t num_requests;         /* R    core.py:98  */
    int32_t area_dimension;       /* dim  core.py:92  */
    int32_t num_racks;            /* L    core.py:93  */
    int32_t racks[WH_MAX_RACKS];  /*      core.py:93  */
    int32_t episode_duration;     /*      core.py:100 */
    int32_t pickup_wait_duration; /*      core.py:101 */
    int32_t max_num_agents;       /*      variants.py:20,36,51 */
    int32_t random_num_agents;    /* 1 = *Train: redraw num_agents in [1,max] on reset (variants.py:69-74) */
} wh_config;

/* Structure-of-arrays state in HBM at the narrowest width holding the reference's value range
 * (reference: int32 everywhere, core.py:153-165). */
typedef struct wh_state {
    int8_t  *agent_pos;     /* [N,R,2]  (x,y)                 core.py:153 */
    int8_t  *agent_tgt;     /* [N,R]    delivery index | -1   core.py:154 */
    int8_t  *pickup_tgt;    /* [N,P]    delivery index | -1   core.py:158 */
    int16_t *pickup_timer;  /* [N,P]    steps left | -1       core.py:159 */
    int32_t *time;          /* [N]      episode_time          core.py:165 */
    int8_t  *num_agents;    /* [N]                            core.py:95  */
    int32_t *episode;       /* [N]      episode counter (RNG stream position), starts at -1 */
    int32_t *acc;           /* [N,4]    per-episode {pickups, deliveries, expired, 0}, 16-byte aligned */
} wh_state;

/* One tensor per key of the observation Dict (core.py:119-148), reference dtypes, [N,R,...]. */
typedef struct wh_obs {
    int32_t *num_agents;             /* [N,R,1]     */
    int32_t *self_position;          /* [N,R,2]     */
    int8_t  *self_availability;      /* [N,R,1]     */
    int32_t *self_delivery_target;   /* [N,R,2]     */
    int32_t *other_positions;        /* [N,R,R-1,2] */
    int8_t  *other_availabilities;   /* [N,R,R-1]   */
    int32_t *other_delivery_targets; /* [N,R,R-1,2] */
    int32_t *requests;               /* [N,R,R,4]   16-byte aligned */
} wh_obs;

/* Render-only mirrors of the state as it was BEFORE the last step (core.py:160-163, 270-272):
 * Warehouse.render(animate=True) interpolates from them (core.py:448-462). */
typedef struct wh_prev {
    int8_t *agent_pos;      /* [N,R,2]  _prev_agent_positions        core.py:161 */
    int8_t *agent_tgt;      /* [N,R]    _prev_agent_delivery_targets core.py:162 */
    int8_t *pickup_tgt;     /* [N,P]    _prev_pickup_point_targets   core.py:163 */
} wh_prev;

#define WH_OBS_STEP  0   /* core.py:371-432 */
#define WH_OBS_RESET 1   /* core.py:224-260 */

#define WH_FLAG_AUTO_RESET 1   /* step: envs whose episode ends are reset in-kernel (native RNG) and
                                  their observation is the first one of the next episode */
#define WH_FLAG_COMPACT_IO 2   /* wh_step: `actions` points to int8 [N,R] and `rewards` to uint8 [N,R]
                                  (values 0/1/2) instead of int32 / float32 — 4x less PCIe traffic for
                                  host-driven loops; semantics unchanged */
#define WH_FLAG_NO_PDL 4       /* launch this step without programmatic stream serialization. By default
                                  the step kernels let the NEXT step launch on the same stream be scheduled
                                  while they drain (it touches no memory before they complete); callers
                                  that put copies between the launches gain nothing from it */

#define WH_FLAG_PER_STEP_OUT 8  /* wh_multi_step: rewards / dones / observations are [T, N, ...] tensors and step t
                                  writes slice t (otherwise: reward sums, last dones, observations overwritten) */

/* wh_multi_step picks its kernel from the launch size; bits 4-6 of `flags` force one (tests, tuning):
 * 0 = automatic, 1 = throughput kernel (256-thread blocks, register-capped), 2 = the same with 64-thread blocks and
 * no register cap, 3 / 4 = warp-specialised kernel with 1 / 2 observation warps per env tile (needs obs and one of
 * the reference variants' geometries, else WH_E_ARG) */
#define WH_FLAG_MULTI_KERNEL(k) (((k) & 7) << 4)

/* stats vector (unsigned 64-bit counters, device memory, WH_NUM_STATS entries):
 * [0] episodes [1] return_sum [2] pickups [3] deliveries [4] expired [5..7] reserved
 * [8+2(n-1)] episodes with n agents, [9+2(n-1)] their return sum  (scripts/train.py:18-23) */

int wh_version(void);
const char *wh_error_string(int code);
int wh_num_pickup_points(const wh_config *cfg);    /* core.py:96 */
int wh_num_delivery_points(const wh_config *cfg);  /* core.py:97 */

/* Warehouse.reset — core.py:167-221 (+ reset observations core.py:224-260 when obs != NULL).
 * Replay mode (agent_pos != NULL): accepted spawn cells [N,R,2], initial request pickup ids and
 * delivery ids [N,R] in draw order, optional num_agents [N] (the *Train redraw). Otherwise the
 * native Philox4x32-10 stream keyed by (seed, env_id0 + e, episode) is used. env_mask [N] (u8,
 * non-zero = reset this env) or NULL for all. */
int wh_reset(const wh_config *cfg, const wh_state *st, int64_t n_envs, int64_t env_id0, uint64_t seed,
             const int8_t *agent_pos, const int8_t *init_pickups, const int8_t *init_targets,
             const int8_t *num_agents, const uint8_t *env_mask, const wh_obs *obs, void *stream);

/* Warehouse.step — core.py:262-368,435-440; fused with the observation build (core.py:371-432)
 * when obs != NULL. actions [N,R] int32; order [N,R] = agent ids in action-dict iteration order,
 * -1 padded (NULL = ascending, core.py:279); spawn_pickups / spawn_targets [N,R] replayed respawn
 * draws, -1 padded (NULL = native RNG); rewards [N,R] f32; dones [N] u8; stats may be NULL.
 * env_mask [N] u8 or NULL: only envs with a non-zero entry are stepped; the others keep their state,
 * observation, rewards and dones untouched (a BaseEnv.send_actions() that covers a subset of the envs). */
int wh_step(const wh_config *cfg, const wh_state *st, int64_t n_envs, int64_t env_id0, uint64_t seed,
            const int32_t *actions, const int32_t *order,
            const int8_t *spawn_pickups, const int8_t *spawn_targets,
            float *rewards, uint8_t *dones, unsigned long long *stats,
            const wh_obs *obs, int flags, const uint8_t *env_mask, void *stream);

/* wh_step with the observations emitted directly in RLlib's flattened float32 layout
 * (see wh_build_obs_flat) by the same kernel — the path a vectorised RLlib sampler consumes.
 * flat_obs [N, R, 9R+1] float32. With WH_FLAG_AUTO_RESET finished envs get their reset observation. */
int wh_step_flat(const wh_config *cfg, const wh_state *st, int64_t n_envs, int64_t env_id0, uint64_t seed,
                 const int32_t *actions, const int32_t *order, float *rewards, uint8_t *dones,
                 unsigned long long *stats, float *flat_obs, int flags, const uint8_t *env_mask, void *stream);

/* Observation build alone — flavour WH_OBS_STEP (core.py:371-432) or WH_OBS_RESET (core.py:224-260). */
int wh_build_obs(const wh_config *cfg, const wh_state *st, int64_t n_envs, int flavour,
                 const wh_obs *obs, void *stream);

/* Observations in RLlib's flattened layout (what RLlib's Dict preprocessor would build from the
 * core.py:119-148 space): float32 out[N, R, 9R+1], keys concatenated in alphabetical order
 * num_agents(1) other_availabilities(R-1) other_delivery_targets(2(R-1)) other_positions(2(R-1))
 * requests(4R) self_availability(1) self_delivery_target(2) self_position(2). */
int wh_build_obs_flat(const wh_config *cfg, const wh_state *st, int64_t n_envs, int flavour,
                      float *out, void *stream);

/* WarehouseRandomGreedySolver.compute_action — solvers.py:27-58 — on observation tensors.
 * rand_threshold = floor(random_action_prob * 2^32). is_random / random_actions [N,R] replay the
 * eps-random branch (solvers.py:44-45); NULL = native RNG keyed by (seed, env, episode, time,
 * agent | 0x80000000 — a stream separate from the env's respawn draws even for equal seeds).
 * actions [N,R] int32 out (-1 for rows >= num_agents). */
int wh_greedy(const wh_config *cfg, const wh_obs *obs, const int8_t *num_agents,
              const int32_t *episode, const int32_t *time, int64_t n_envs, int64_t env_id0,
              uint64_t seed, uint64_t rand_threshold, const uint8_t *is_random,
              const int32_t *random_actions, int32_t *actions, void *stream);

/* run.py:42-62 loop body for all envs in ONE kernel: greedy solver evaluated from the state held
 * in registers (identical to solving on the previous observation, incl. the reset-flavour obs at
 * time 0) -> step -> observation build. actions_out [N,R] may be NULL. */
int wh_greedy_step(const wh_config *cfg, const wh_state *st, int64_t n_envs, int64_t env_id0,
                   uint64_t seed, uint64_t solver_seed, uint64_t rand_threshold,
                   int32_t *actions_out, float *rewards, uint8_t *dones,
                   unsigned long long *stats, const wh_obs *obs, int flags, void *stream);

/* `n_steps` iterations of baseline/run.py:42-62 (greedy solver -> step) in ONE launch, for evaluating
 * the baseline at scale: the state stays in registers between the steps and no observation is written
 * (call wh_build_obs afterwards if one is needed). Leaves state, statistics and dones exactly as n_steps
 * wh_greedy_step launches would; reward_sums [N,R] = each agent's reward summed over these steps. */
int wh_greedy_rollout(const wh_config *cfg, const wh_state *st, int64_t n_envs, int64_t env_id0,
                      uint64_t seed, uint64_t solver_seed, uint64_t rand_threshold, int n_steps,
                      float *reward_sums, uint8_t *dones, unsigned long long *stats, int flags, void *stream);

/* `n_steps` consecutive Warehouse.step calls (core.py:262-442) in ONE launch, the state held in registers
 * between the steps; every step writes what a single wh_step / wh_greedy_step launch writes. For batch
 * sizes where one step is shorter than a kernel launch (BASELINE configs[1]: 4 096 Small envs) and for
 * open-loop replays (recorded / random action tensors, the greedy baseline).
 *   actions  int32 [n_steps, N, R] open-loop actions (-1 = absent), or NULL = the in-kernel greedy solver
 *            (solvers.py:27-58; solver_seed / rand_threshold as in wh_greedy_step);
 *   flags    WH_FLAG_AUTO_RESET, WH_FLAG_PER_STEP_OUT, WH_FLAG_MULTI_KERNEL(k);
 *   rewards  f32 [n_steps,N,R] / dones u8 [n_steps,N] with WH_FLAG_PER_STEP_OUT, else [N,R] per-agent
 *            reward SUMS over the steps and [N] the last step's dones;
 *   obs      NULL = no observations; otherwise written EVERY step: into slice t of [n_steps,N,...] tensors
 *            with WH_FLAG_PER_STEP_OUT, else over the same [N,...] tensors (what n_steps launches leave).
 * Moves are resolved in ascending agent order (core.py:279 default); native RNG only. */
int wh_multi_step(const wh_config *cfg, const wh_state *st, int64_t n_envs, int64_t env_id0, uint64_t seed,
                  int n_steps, const int32_t *actions, uint64_t solver_seed, uint64_t rand_threshold,
                  float *rewards, uint8_t *dones, unsigned long long *stats, const wh_obs *obs, int flags,
                  void *stream);

/* core.py:270-272 — keep the render-only `_prev_*` mirrors: copies agent positions, agent delivery
 * targets and pickup-point targets into `prev` on `stream` (device-to-device, no synchronisation). A caller
 * that wants Warehouse.render(animate=True) issues it right before wh_step; nobody else pays for it. */
int wh_save_prev(const wh_config *cfg, const wh_state *st, const wh_prev *prev, int64_t n_envs, void *stream);

/* End-of-rollout reduction of the episode statistics over NVLink: in-place ncclAllReduce(sum) of the
 * WH_NUM_STATS uint64 counters on `stream`. `nccl_comm` is a ncclComm_t created by the caller with
 * the NCCL already loaded in the process (the symbol is resolved at run time with dlsym, so the
 * library has no link-time NCCL dependency). This is the only collective on the path. */
int wh_stats_allreduce(unsigned long long *stats, void *nccl_comm, void *stream);

/* ---- Layer 2: host-buffer environment handle -------------------------------------------- */
typedef struct wh_env wh_env;

/* Allocates device state + observation tensors for n_envs on `device` and `n_chunks` streams (the
 * env batch is processed as n_chunks copy -> kernel -> copy pipelines). Observations always stay
 * resident in HBM (wh_env_obs_ptrs); wh_env_step_host copies them out only when given obs_host.
 * n_chunks selects how a step's host I/O crosses PCIe:
 *   k > 0   copy pipeline: k chunks of [cudaMemcpyAsync H2D actions -> kernel -> cudaMemcpyAsync D2H];
 *   <= 0    direct: ONE kernel over all envs reads the actions from, and writes the rewards to, the caller's
 *           page-locked host buffers itself (mapped / UVA access: no copy to wait for before or after the
 *           kernel, no cross-engine dependency); the one-byte-per-env dones stay a device tensor followed by
 *           one small copy (single-byte stores over PCIe cost more than the whole rest of the step). Needs
 *           cudaHostAlloc / cudaHostRegister / torch pin_memory buffers, else falls back to k = 1.
 * On failure nothing is left allocated and *out is NULL. */
int wh_env_create(const wh_config *cfg, int64_t n_envs, int device, int64_t env_id0, uint64_t seed,
                  int n_chunks, wh_env **out);
void wh_env_destroy(wh_env *env);
int wh_env_reset(wh_env *env);
/* HOST buffers: actions [N,R] int32 in; rewards [N,R] f32 and dones [N] u8 out. Copies and
 * kernels are pipelined per env chunk; returns after everything has landed in the host buffers.
 * obs_host (8 host pointers in wh_obs order, or NULL) additionally copies the observations out. */
int wh_env_step_host(wh_env *env, const int32_t *actions, float *rewards, uint8_t *dones,
                     const wh_obs *obs_host);
/* Same with WH_FLAG_COMPACT_IO dtypes on the wire: int8 actions in, uint8 rewards + uint8 dones out. */
int wh_env_step_host_compact(wh_env *env, const int8_t *actions, uint8_t *rewards, uint8_t *dones);
/* greedy-policy variant: the solver runs on device, nothing goes H2D; rewards/dones come back. */
int wh_env_greedy_step_host(wh_env *env, float *rewards, uint8_t *dones);
int wh_env_obs_ptrs(wh_env *env, wh_obs *out);      /* device pointers of the resident obs */
int wh_env_state_ptrs(wh_env *env, wh_state *out);  /* device pointers of the resident state */
int wh_env_stats_host(wh_env *env, unsigned long long *stats_out /* [WH_NUM_STATS] */);
int64_t wh_env_launch_count(wh_env *env);           /* kernels launched so far through this handle */

#ifdef __cplusplus
}
#endif
#endif /* WH_B200_H */
