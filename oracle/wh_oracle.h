/*
 * wh_oracle.h — CPU ORACLE for the warehouse hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the reference algorithm (ffahleraz/rllib-warehouse):
 *   warehouse/core.py:167-260  (reset)      warehouse/core.py:262-442 (step + observations)
 *   warehouse/variants.py:19-98 (constants) baseline/solvers.py:27-58 (greedy solver)
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this. The product path (rllib_warehouse_b200) never links or calls it.
 *
 * Parity pinning: the reference ships no tests or golden vectors, so this oracle
 * is pinned against fixtures produced by EXECUTING the unmodified reference
 * (oracle/make_golden.py -> the .npz fixtures under tests/golden; checked by tests/test_oracle_golden.py) and, in
 * the build container, against the live reference (tests/test_oracle_vs_reference.py).
 *
 * All state is held at the reference's own width (int32) so that every array can be compared
 * one-to-one with the reference's numpy arrays. Agent rows are padded to R = num_requests
 * (rows >= num_agents hold -1 and never act).
 */
#ifndef WH_ORACLE_H
#define WH_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WHO_MAX_RACKS 8

typedef struct who_config {
    int32_t num_requests;        /* R   core.py:98  */
    int32_t area_dimension;      /* dim core.py:92  */
    int32_t num_racks;           /* L = len(pickup_racks_arrangement) core.py:93 */
    int32_t racks[WHO_MAX_RACKS];
    int32_t episode_duration;    /* core.py:100 */
    int32_t pickup_wait_duration;/* core.py:101 */
    int32_t max_num_agents;      /* variants.py max_num_agents (== R for the six variants) */
    int32_t random_num_agents;   /* 1 = *Train variants: redraw num_agents on reset (variants.py:69-74) */
} who_config;

/* Per-env state, env-major, int32 (reference dtypes, core.py:153-165). */
typedef struct who_state {
    int32_t *agent_pos;      /* [N,R,2]  core.py:153  (-1 rows for agents >= num_agents) */
    int32_t *agent_tgt;      /* [N,R]    core.py:154  delivery-point index or -1 */
    int32_t *pickup_tgt;     /* [N,P]    core.py:158 */
    int32_t *pickup_timer;   /* [N,P]    core.py:159 */
    int32_t *time;           /* [N]      core.py:165 */
    int32_t *num_agents;     /* [N]      core.py:95  */
    int32_t *episode;        /* [N]      episode counter (RNG stream position; not in the reference) */
    int32_t *acc;            /* [N,4]    per-episode {pickups, deliveries, expired, 0} */
} who_state;

/* Observation tensors, one per key of core.py:119-148, env-major [N,R,...]. */
typedef struct who_obs {
    int32_t *num_agents;             /* [N,R,1]      */
    int32_t *self_position;          /* [N,R,2]      */
    int8_t  *self_availability;      /* [N,R,1]      */
    int32_t *self_delivery_target;   /* [N,R,2]      */
    int32_t *other_positions;        /* [N,R,R-1,2]  */
    int8_t  *other_availabilities;   /* [N,R,R-1]    */
    int32_t *other_delivery_targets; /* [N,R,R-1,2]  */
    int32_t *requests;               /* [N,R,R,4]    */
} who_obs;

#define WHO_NUM_STATS 80   /* see who_stats_index() in wh_oracle.c */

int who_num_pickup_points(const who_config *cfg);     /* 4*L*L   core.py:96 */
int who_num_delivery_points(const who_config *cfg);   /* 4*(dim-4) core.py:97 */
void who_pickup_cell(const who_config *cfg, int p, int *x, int *y);     /* core.py:171-175 */
void who_delivery_cell(const who_config *cfg, int d, int *x, int *y);   /* core.py:178-188 */

/* reset (core.py:167-221). Replay mode when agent_pos != NULL: agent cells [N,R,2] (rows < A used),
 * init_pickups/init_targets [N,R], num_agents_in [N] (or NULL to keep state's). Otherwise the
 * native counter-based RNG (Philox4x32-10 keyed by seed, env id = env_id0 + e) is used.
 * env_mask (u8 [N]) selects envs, NULL = all. */
int who_reset(const who_config *cfg, who_state *st, int64_t n_envs, int64_t env_id0, uint64_t seed,
              const int32_t *agent_pos, const int32_t *init_pickups, const int32_t *init_targets,
              const int32_t *num_agents_in, const uint8_t *env_mask);

/* step without observations (core.py:262-368, 435-440). actions [N,R] (-1 = absent from the
 * action dict); order [N,R] = agent ids in action-dict iteration order, -1 padded, or NULL for
 * ascending; spawn_pickups/spawn_targets [N,R] replayed draws (-1 padded) or NULL for native RNG.
 * stats: int64[WHO_NUM_STATS] accumulators or NULL. */
int who_step(const who_config *cfg, who_state *st, int64_t n_envs, int64_t env_id0, uint64_t seed,
             const int32_t *actions, const int32_t *order,
             const int32_t *spawn_pickups, const int32_t *spawn_targets,
             float *rewards, uint8_t *dones, int64_t *stats);

/* observation build (core.py:224-260 reset flavour = 1; core.py:371-432 step flavour = 0). */
int who_build_obs(const who_config *cfg, const who_state *st, int64_t n_envs, int flavour,
                  who_obs *obs);

/* greedy solver (solvers.py:27-58) on observation tensors. rand_threshold = floor(p * 2^32);
 * is_random/random_actions [N,R] replay the eps-random branch (solvers.py:44-45) or NULL for the
 * native RNG keyed by (seed, env, episode, time). actions out [N,R] (-1 for rows >= num_agents). */
int who_greedy(const who_config *cfg, const who_obs *obs, const int32_t *num_agents,
               const int32_t *episode, const int32_t *time,
               int64_t n_envs, int64_t env_id0, uint64_t seed, uint64_t rand_threshold,
               const uint8_t *is_random, const int32_t *random_actions, int32_t *actions);

/* Multi-threaded rollout used as the CPU baseline: n_steps of {policy} -> step -> build_obs over all
 * envs with `n_threads` pthreads, envs partitioned contiguously (no synchronisation between
 * threads: envs are independent). policy 0: `actions` [n_action_sets,N,R], step s applies set
 * s % n_action_sets (random-action workload); policy 1: the greedy solver on the previous observations, written to
 * actions_scratch [N,R]. auto_reset: finished envs are reset with the native RNG and get their
 * reset observation. Returns the total number of agent-steps executed. */
int64_t who_rollout(const who_config *cfg, who_state *st, who_obs *obs, int64_t n_envs,
                    int64_t env_id0, uint64_t seed, int policy, const int32_t *actions,
                    int32_t *actions_scratch, float *rewards, uint8_t *dones, int64_t *stats,
                    int n_steps, int n_threads, int auto_reset, int n_action_sets);

/* Raw Philox4x32-10 block, exported so tests can check the CUDA RNG directly. */
void who_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint64_t seed, uint32_t out[4]);

#ifdef __cplusplus
}
#endif
#endif
