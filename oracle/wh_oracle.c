/*
 * wh_oracle.c — CPU ORACLE for the warehouse hot path.  TEST INFRASTRUCTURE ONLY.
 * See wh_oracle.h for scope, pinning and who may load this.
 *
 * The implementation deliberately follows the reference LITERALLY (occupancy grid + explicit
 * invalid-move list, array scans in index order) rather than the bit-mask / ballot
 * reformulation used by the CUDA kernels, so that agreement between the two is evidence and
 * not tautology. Citations are to /root/reference/.
 */
#include "wh_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define MAX_DIM 128
#define MAX_R 32
#define MAX_P 64
#define MAX_D 64

/* ------------------------------------------------------------------------------------------ */
/* static tables                                                                               */
/* ------------------------------------------------------------------------------------------ */

int who_num_pickup_points(const who_config *cfg) { return 4 * cfg->num_racks * cfg->num_racks; }
int who_num_delivery_points(const who_config *cfg) { return 4 * (cfg->area_dimension - 4); }

/* core.py:171-175: for x in racks: for y in racks: (x-1,y-1),(x,y-1),(x-1,y),(x,y) */
void who_pickup_cell(const who_config *cfg, int p, int *x, int *y) {
    static const int ox[4] = {-1, 0, -1, 0}, oy[4] = {-1, -1, 0, 0};
    int L = cfg->num_racks;
    int rx = cfg->racks[p / (4 * L)], ry = cfg->racks[(p / 4) % L], c = p % 4;
    *x = rx + ox[c];
    *y = ry + oy[c];
}

/* core.py:178-188: for val in range(2, dim-2): (val,0),(0,val),(val,dim-1),(dim-1,val) */
void who_delivery_cell(const who_config *cfg, int d, int *x, int *y) {
    int v = 2 + d / 4, dim = cfg->area_dimension;
    switch (d % 4) {
    case 0: *x = v; *y = 0; break;
    case 1: *x = 0; *y = v; break;
    case 2: *x = v; *y = dim - 1; break;
    default: *x = dim - 1; *y = v; break;
    }
}

static int is_pickup_cell(const who_config *cfg, int x, int y) {
    int P = who_num_pickup_points(cfg);
    for (int p = 0; p < P; ++p) {
        int px, py;
        who_pickup_cell(cfg, p, &px, &py);
        if (px == x && py == y) return 1;
    }
    return 0;
}

static int check_cfg(const who_config *cfg) {
    int P = who_num_pickup_points(cfg), D = who_num_delivery_points(cfg);
    if (cfg->num_requests < 1 || cfg->num_requests > MAX_R) return 1;
    if (cfg->area_dimension < 5 || cfg->area_dimension > MAX_DIM) return 1;
    if (cfg->num_racks < 1 || cfg->num_racks > WHO_MAX_RACKS) return 1;
    if (P > MAX_P || D > MAX_D || cfg->num_requests > P || cfg->num_requests > D) return 1;
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* native counter-based RNG (NOT part of the reference: the reference uses the process-global   */
/* numpy MT19937 stream, which is replayed through the *_replay inputs instead)     */
/* ------------------------------------------------------------------------------------------ */

void who_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint64_t seed, uint32_t out[4]) {
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static uint32_t bounded(uint32_t u, uint32_t n) { return (uint32_t)(((uint64_t)u * n) >> 32); }

/* index of the n-th (0-based, ascending) set bit */
static int nth_set_bit(uint64_t mask, int n) {
    for (int b = 0; b < 64; ++b)
        if ((mask >> b) & 1) { if (n == 0) return b; --n; }
    return -1;
}

#define CTR_NUM_AGENTS 0xFFFFFFFFu
#define CTR_SPAWN_AGENT 0xF0000000u
#define CTR_INIT_REQUESTS 0xE0000000u

/* ------------------------------------------------------------------------------------------ */
/* stats                                                                                       */
/* ------------------------------------------------------------------------------------------ */
/* [0] episodes  [1] return_sum  [2] pickups  [3] deliveries  [4] expired  [5..7] reserved
 * [8+2(n-1)] episodes with n agents   [9+2(n-1)] return_sum with n agents   (n = 1..36)
 * mirrors scripts/train.py:18-23 (avg_agent_reward_all / avg_agent_reward_{n}). */
static void stats_flush(int64_t *stats, const int32_t *acc, int A) {
    int64_t ret = (int64_t)acc[0] + acc[1];
    stats[0] += 1; stats[1] += ret; stats[2] += acc[0]; stats[3] += acc[1]; stats[4] += acc[2];
    stats[8 + 2 * (A - 1)] += 1;
    stats[9 + 2 * (A - 1)] += ret;
}

/* ------------------------------------------------------------------------------------------ */
/* reset — core.py:167-221                                                                     */
/* ------------------------------------------------------------------------------------------ */
static void reset_one(const who_config *cfg, who_state *st, int64_t e, int64_t env_id, uint64_t seed,
                      const int32_t *agent_pos, const int32_t *init_pickups,
                      const int32_t *init_targets, const int32_t *num_agents_in) {
    const int R = cfg->num_requests, P = who_num_pickup_points(cfg), D = who_num_delivery_points(cfg);
    const int dim = cfg->area_dimension;
    int32_t *pos = st->agent_pos + e * R * 2, *tgt = st->agent_tgt + e * R;
    int32_t *ptgt = st->pickup_tgt + e * P, *ptim = st->pickup_timer + e * P;
    uint32_t rnd[4];
    st->episode[e] += 1;
    const uint32_t ep = (uint32_t)st->episode[e], eid = (uint32_t)env_id;
    st->time[e] = 0;                                                         /* core.py:168 */
    int A = st->num_agents[e];
    if (agent_pos) {
        if (num_agents_in) A = num_agents_in[e];
    } else if (cfg->random_num_agents) {                                     /* variants.py:70,74 */
        who_philox(eid, ep, CTR_NUM_AGENTS, 0, seed, rnd);
        A = 1 + (int)bounded(rnd[0], (uint32_t)cfg->max_num_agents);
    }
    st->num_agents[e] = A;
    for (int a = 0; a < R; ++a) {
        int x = -1, y = -1;
        if (a < A) {
            if (agent_pos) { x = agent_pos[(e * R + a) * 2]; y = agent_pos[(e * R + a) * 2 + 1]; }
            else {
                /* core.py:192-201: rejection-sample randint(1, dim-1)^2 until not a pickup cell;
                 * other agents are NOT checked (co-location possible) */
                for (uint32_t j = 0;; ++j) {
                    who_philox(eid, ep, CTR_SPAWN_AGENT + j, (uint32_t)a, seed, rnd);
                    x = 1 + (int)bounded(rnd[0], (uint32_t)(dim - 2));
                    y = 1 + (int)bounded(rnd[1], (uint32_t)(dim - 2));
                    if (!is_pickup_cell(cfg, x, y)) break;
                }
            }
        }
        pos[2 * a] = x; pos[2 * a + 1] = y;
        tgt[a] = -1;                                                          /* core.py:204 */
    }
    for (int p = 0; p < P; ++p) { ptgt[p] = -1; ptim[p] = -1; }               /* core.py:210-211 */
    /* core.py:215-221: R distinct pickup points, R distinct delivery points, paired in draw order */
    uint64_t inactive = (P == 64) ? ~0ull : ((1ull << P) - 1);
    uint64_t avail_d = (D == 64) ? ~0ull : ((1ull << D) - 1);
    for (int i = 0; i < R; ++i) {
        int p, d;
        if (agent_pos) { p = init_pickups[e * R + i]; d = init_targets[e * R + i]; }
        else {
            who_philox(eid, ep, CTR_INIT_REQUESTS, (uint32_t)i, seed, rnd);
            p = nth_set_bit(inactive, (int)bounded(rnd[0], (uint32_t)(P - i)));
            d = nth_set_bit(avail_d, (int)bounded(rnd[1], (uint32_t)(D - i)));
            inactive &= ~(1ull << p);
            avail_d &= ~(1ull << d);
        }
        if (p < 0) continue;
        ptgt[p] = d;
        ptim[p] = cfg->pickup_wait_duration;
    }
    st->acc[4 * e] = st->acc[4 * e + 1] = st->acc[4 * e + 2] = st->acc[4 * e + 3] = 0;
}

int who_reset(const who_config *cfg, who_state *st, int64_t n_envs, int64_t env_id0, uint64_t seed,
              const int32_t *agent_pos, const int32_t *init_pickups, const int32_t *init_targets,
              const int32_t *num_agents_in, const uint8_t *env_mask) {
    if (check_cfg(cfg)) return 1;
    for (int64_t e = 0; e < n_envs; ++e)
        if (!env_mask || env_mask[e])
            reset_one(cfg, st, e, env_id0 + e, seed, agent_pos, init_pickups, init_targets, num_agents_in);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* step — core.py:262-368, 435-440                                                             */
/* ------------------------------------------------------------------------------------------ */
typedef struct { int a, b, c, d; } move4;

static void step_one(const who_config *cfg, who_state *st, int64_t e, int64_t env_id, uint64_t seed,
                     const int32_t *actions, const int32_t *order,
                     const int32_t *spawn_pickups, const int32_t *spawn_targets,
                     float *rewards, uint8_t *dones, int64_t *stats) {
    const int R = cfg->num_requests, P = who_num_pickup_points(cfg), D = who_num_delivery_points(cfg);
    const int dim = cfg->area_dimension;
    const int A = st->num_agents[e];
    int32_t *pos = st->agent_pos + e * R * 2, *tgt = st->agent_tgt + e * R;
    int32_t *ptgt = st->pickup_tgt + e * P, *ptim = st->pickup_timer + e * P;
    int32_t *acc = st->acc + 4 * e;

    st->time[e] += 1;                                                         /* core.py:267 */

    /* --- core.py:275-277: occupancy grid marks ALL agents, acting or not -------------------- */
    static __thread uint8_t occ[MAX_DIM][MAX_DIM];
    for (int x = 0; x < dim; ++x) memset(occ[x], 0, (size_t)dim);
    for (int a = 0; a < A; ++a) occ[pos[2 * a]][pos[2 * a + 1]] = 1;
    move4 invalid[3 * MAX_R];
    int n_invalid = 0;

    /* --- core.py:279-300: sequential moves in action-dict iteration order ------------------- */
    for (int t = 0; t < R; ++t) {
        int idx = order ? order[e * R + t] : t;
        if (order && idx < 0) break;
        if (idx < 0 || idx >= A) continue;
        int action = actions[e * R + idx];
        if (action < 0) continue;                       /* agent absent from the dict */
        int px = pos[2 * idx], py = pos[2 * idx + 1];
        int x = px + (action / 3 - 1), y = py + (action % 3 - 1);   /* MOVES, core.py:38,282 */
        if (!(0 <= x && x < dim)) x = px;                           /* core.py:284-287 */
        if (!(0 <= y && y < dim)) y = py;
        int bad = occ[x][y];
        for (int i = 0; i < n_invalid && !bad; ++i)
            bad = invalid[i].a == px && invalid[i].b == py && invalid[i].c == x && invalid[i].d == y;
        if (bad) continue;                                          /* core.py:289 */
        occ[px][py] = 0;                                            /* core.py:290-291 */
        occ[x][y] = 1;
        invalid[n_invalid++] = (move4){x, y, px, py};               /* core.py:294 */
        if (x != px && y != py) {                                   /* core.py:295-297 */
            invalid[n_invalid++] = (move4){x, py, px, y};
            invalid[n_invalid++] = (move4){px, y, x, py};
        }
        pos[2 * idx] = x; pos[2 * idx + 1] = y;                     /* core.py:299-300 */
    }

    /* --- core.py:303-306: expiry ------------------------------------------------------------ */
    int n_expired = 0;
    for (int p = 0; p < P; ++p) if (ptgt[p] > -1) ptim[p] -= 1;
    for (int p = 0; p < P; ++p) if (ptim[p] == 0) { ptgt[p] = -1; ptim[p] = -1; ++n_expired; }

    /* --- core.py:309-335: pickups ----------------------------------------------------------- */
    int cand[MAX_R], picks[MAX_R];
    for (int a = 0; a < A; ++a) {
        cand[a] = -1;                                   /* argmax of the collision row = FIRST match */
        for (int p = 0; p < P && cand[a] < 0; ++p) {
            int cx, cy;
            who_pickup_cell(cfg, p, &cx, &cy);
            if (cx == pos[2 * a] && cy == pos[2 * a + 1]) cand[a] = p;
        }
        /* the mask is evaluated for all agents against the PRE-assignment targets (core.py:320-324) */
        picks[a] = cand[a] >= 0 && tgt[a] == -1 && ptgt[cand[a]] > -1;
    }
    int n_pick = 0;
    for (int a = 0; a < A; ++a) {
        rewards[e * R + a] = 0.0f;                                   /* core.py:334 */
        if (picks[a]) { tgt[a] = ptgt[cand[a]]; }                    /* core.py:327-329 (gather first) */
    }
    for (int a = 0; a < A; ++a)
        if (picks[a]) { ptgt[cand[a]] = -1; ptim[cand[a]] = -1;      /* core.py:330-331 */
                        rewards[e * R + a] += 1.0f; ++n_pick; }      /* core.py:335, PICKUP_REWARD */
    for (int a = A; a < R; ++a) rewards[e * R + a] = 0.0f;

    /* --- core.py:338-351: respawn so that exactly R requests are active ---------------------- */
    int inact[MAX_P], n_inact = 0;
    for (int p = 0; p < P; ++p) if (ptgt[p] == -1) inact[n_inact++] = p;
    int k = R - P + n_inact;
    if (spawn_pickups) {
        for (int i = 0; i < R; ++i) {
            int p = spawn_pickups[e * R + i];
            if (p < 0) break;
            ptim[p] = cfg->pickup_wait_duration;                     /* core.py:344 */
            ptgt[p] = spawn_targets[e * R + i];                      /* core.py:351 */
        }
    } else {
        uint64_t inactive = 0, avail_d = (D == 64) ? ~0ull : ((1ull << D) - 1);
        for (int i = 0; i < n_inact; ++i) inactive |= 1ull << inact[i];
        const uint32_t ep = (uint32_t)st->episode[e], eid = (uint32_t)env_id;
        for (int i = 0; i < k; ++i) {
            uint32_t rnd[4];
            who_philox(eid, ep, (uint32_t)st->time[e], (uint32_t)i, seed, rnd);
            int p = nth_set_bit(inactive, (int)bounded(rnd[0], (uint32_t)(n_inact - i)));
            int d = nth_set_bit(avail_d, (int)bounded(rnd[1], (uint32_t)(D - i)));
            inactive &= ~(1ull << p);
            avail_d &= ~(1ull << d);
            ptim[p] = cfg->pickup_wait_duration;
            ptgt[p] = d;
        }
    }

    /* --- core.py:354-368: deliveries (agents that picked up THIS step are already delivering) - */
    int n_deliv = 0;
    for (int a = 0; a < A; ++a) {
        if (tgt[a] > -1) {
            int dx, dy;
            who_delivery_cell(cfg, tgt[a], &dx, &dy);
            if (dx == pos[2 * a] && dy == pos[2 * a + 1]) {
                tgt[a] = -1;
                rewards[e * R + a] += 1.0f;                           /* DELIVERY_REWARD */
                ++n_deliv;
            }
        }
    }

    /* --- core.py:438-440 ---------------------------------------------------------------------- */
    dones[e] = st->time[e] >= cfg->episode_duration;

    acc[0] += n_pick; acc[1] += n_deliv; acc[2] += n_expired;
    if (stats && st->time[e] == cfg->episode_duration) stats_flush(stats, acc, A);
}

int who_step(const who_config *cfg, who_state *st, int64_t n_envs, int64_t env_id0, uint64_t seed,
             const int32_t *actions, const int32_t *order,
             const int32_t *spawn_pickups, const int32_t *spawn_targets,
             float *rewards, uint8_t *dones, int64_t *stats) {
    if (check_cfg(cfg)) return 1;
    for (int64_t e = 0; e < n_envs; ++e)
        step_one(cfg, st, e, env_id0 + e, seed, actions, order, spawn_pickups, spawn_targets,
                 rewards, dones, stats);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* observations — core.py:224-260 (reset flavour), core.py:371-432 (step flavour)              */
/* ------------------------------------------------------------------------------------------ */
static void obs_one(const who_config *cfg, const who_state *st, int64_t e, int flavour, who_obs *o) {
    const int R = cfg->num_requests, P = who_num_pickup_points(cfg);
    const int null_pos = cfg->area_dimension / 2;                    /* core.py:107 */
    const int A = st->num_agents[e];
    const int32_t *pos = st->agent_pos + e * R * 2, *tgt = st->agent_tgt + e * R;
    const int32_t *ptgt = st->pickup_tgt + e * P;
    int32_t ppos[MAX_R][2], tpos[MAX_R][2], req[MAX_R][4];
    int8_t avail[MAX_R];
    for (int r = 0; r < R; ++r) {                     /* padded [R] tables, core.py:372-407 */
        int real = r < A;
        ppos[r][0] = real ? pos[2 * r] : null_pos;
        ppos[r][1] = real ? pos[2 * r + 1] : null_pos;
        int delivering = real && tgt[r] > -1;
        /* reset flavour: every availability 0, every delivery-target null (core.py:233-236) */
        avail[r] = (flavour == 0 && real && !delivering) ? 1 : 0;   /* core.py:383-384 */
        tpos[r][0] = tpos[r][1] = null_pos;
        if (flavour == 0 && delivering) who_delivery_cell(cfg, tgt[r], &tpos[r][0], &tpos[r][1]);
    }
    int n_req = 0;                                    /* core.py:409-418: ascending pickup index */
    for (int p = 0; p < P && n_req < R; ++p)
        if (ptgt[p] > -1) {
            who_pickup_cell(cfg, p, &req[n_req][0], &req[n_req][1]);
            who_delivery_cell(cfg, ptgt[p], &req[n_req][2], &req[n_req][3]);
            ++n_req;
        }
    for (; n_req < R; ++n_req)                        /* unreachable from reset/step (invariant: R active) */
        req[n_req][0] = req[n_req][1] = req[n_req][2] = req[n_req][3] = null_pos;

    for (int i = 0; i < R; ++i) {                     /* rows >= A follow the same formula on padding */
        int64_t row = e * R + i;
        o->num_agents[row] = A;
        o->self_position[row * 2] = ppos[i][0];
        o->self_position[row * 2 + 1] = ppos[i][1];
        o->self_availability[row] = avail[i];
        o->self_delivery_target[row * 2] = tpos[i][0];
        o->self_delivery_target[row * 2 + 1] = tpos[i][1];
        /* core.py:426-428: other_positions / other_availabilities delete row i;
         * other_delivery_targets deletes row i at reset (core.py:256) but ALWAYS ROW 1 in step (core.py:428) */
        int del_t = flavour == 0 ? 1 : i;
        for (int o_ = 0, s = 0, s2 = 0; o_ < R - 1; ++o_, ++s, ++s2) {
            if (s == i) ++s;
            if (s2 == del_t) ++s2;
            int64_t k = row * (R - 1) + o_;
            o->other_positions[k * 2] = ppos[s][0];
            o->other_positions[k * 2 + 1] = ppos[s][1];
            o->other_availabilities[k] = avail[s];
            o->other_delivery_targets[k * 2] = tpos[s2][0];
            o->other_delivery_targets[k * 2 + 1] = tpos[s2][1];
        }
        memcpy(o->requests + row * R * 4, req, sizeof(int32_t) * 4 * (size_t)R);
    }
}

int who_build_obs(const who_config *cfg, const who_state *st, int64_t n_envs, int flavour, who_obs *obs) {
    if (check_cfg(cfg)) return 1;
    for (int64_t e = 0; e < n_envs; ++e) obs_one(cfg, st, e, flavour, obs);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* greedy solver — baseline/solvers.py:27-58                                                   */
/* ------------------------------------------------------------------------------------------ */
static int clip1(int v) { return v < -1 ? -1 : (v > 1 ? 1 : v); }

static void greedy_one(const who_config *cfg, const who_obs *o, const int32_t *num_agents,
                       const int32_t *episode, const int32_t *time, int64_t e, int64_t env_id,
                       uint64_t seed, uint64_t thr, const uint8_t *is_random,
                       const int32_t *random_actions, int32_t *actions) {
    const int R = cfg->num_requests, A = num_agents[e];
    for (int i = 0; i < R; ++i) {
        int64_t row = e * R + i;
        if (i >= A) { actions[row] = -1; continue; }
        int sx = o->self_position[row * 2], sy = o->self_position[row * 2 + 1];
        int tx, ty;
        if (o->self_availability[row] == 0) {                       /* solvers.py:33-34 */
            tx = o->self_delivery_target[row * 2];
            ty = o->self_delivery_target[row * 2 + 1];
        } else {                                                    /* solvers.py:53-58: L1 argmin, first min */
            const int32_t *rq = o->requests + row * R * 4;
            int best = 0, bestd = 1 << 30;
            for (int j = 0; j < R; ++j) {
                int d = abs(sx - rq[4 * j]) + abs(sy - rq[4 * j + 1]);
                if (d < bestd) { bestd = d; best = j; }
            }
            tx = rq[4 * best]; ty = rq[4 * best + 1];
        }
        int action = (clip1(tx - sx) + 1) * 3 + (clip1(ty - sy) + 1);   /* solvers.py:41,47-49 */
        if (is_random) {                                            /* solvers.py:44-45, replayed */
            if (is_random[row]) action = random_actions[row];
        } else if (thr) {
            uint32_t rnd[4];
            /* solver stream: agent index tagged with bit 31, so it never shares a counter with the env's
             * respawn draws of the same (env, episode, time) even when solver_seed == seed */
            who_philox((uint32_t)env_id, (uint32_t)episode[e], (uint32_t)time[e], (uint32_t)i | 0x80000000u, seed, rnd);
            if ((uint64_t)rnd[0] < thr) action = (int)bounded(rnd[1], 9);
        }
        actions[row] = action;
    }
}

int who_greedy(const who_config *cfg, const who_obs *obs, const int32_t *num_agents,
               const int32_t *episode, const int32_t *time,
               int64_t n_envs, int64_t env_id0, uint64_t seed, uint64_t rand_threshold,
               const uint8_t *is_random, const int32_t *random_actions, int32_t *actions) {
    if (check_cfg(cfg)) return 1;
    for (int64_t e = 0; e < n_envs; ++e)
        greedy_one(cfg, obs, num_agents, episode, time, e, env_id0 + e, seed, rand_threshold,
                   is_random, random_actions, actions);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* multi-threaded rollout (CPU baseline leg of bench.py)                                       */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    const who_config *cfg; who_state *st; who_obs *obs;
    int64_t e0, e1, env_id0, n_envs; uint64_t seed; int policy, n_action_sets; const int32_t *actions;
    int32_t *scratch; float *rewards; uint8_t *dones; int64_t stats[WHO_NUM_STATS];
    int n_steps, auto_reset; int64_t agent_steps;
} job_t;

static void *rollout_thread(void *arg) {
    job_t *j = (job_t *)arg;
    for (int s = 0; s < j->n_steps; ++s) {
        for (int64_t e = j->e0; e < j->e1; ++e) {
            const int32_t *act = j->actions + (int64_t)(s % j->n_action_sets) * j->n_envs * j->cfg->num_requests;
            if (j->policy == 1) {
                greedy_one(j->cfg, j->obs, j->st->num_agents, j->st->episode, j->st->time, e,
                           j->env_id0 + e, j->seed ^ 0x5EEDull, 0, NULL, NULL, j->scratch);
                act = j->scratch;
            }
            step_one(j->cfg, j->st, e, j->env_id0 + e, j->seed, act, NULL, NULL, NULL,
                     j->rewards, j->dones, j->stats);
            j->agent_steps += j->st->num_agents[e];
            if (j->auto_reset && j->dones[e]) {
                reset_one(j->cfg, j->st, e, j->env_id0 + e, j->seed, NULL, NULL, NULL, NULL);
                obs_one(j->cfg, j->st, e, 1, j->obs);
            } else {
                obs_one(j->cfg, j->st, e, 0, j->obs);
            }
        }
    }
    return NULL;
}

int64_t who_rollout(const who_config *cfg, who_state *st, who_obs *obs, int64_t n_envs,
                    int64_t env_id0, uint64_t seed, int policy, const int32_t *actions,
                    int32_t *actions_scratch, float *rewards, uint8_t *dones, int64_t *stats,
                    int n_steps, int n_threads, int auto_reset, int n_action_sets) {
    if (check_cfg(cfg) || n_threads < 1) return -1;
    if (n_threads > n_envs) n_threads = (int)n_envs;
    job_t *jobs = (job_t *)calloc((size_t)n_threads, sizeof(job_t));
    pthread_t *tid = (pthread_t *)calloc((size_t)n_threads, sizeof(pthread_t));
    for (int t = 0; t < n_threads; ++t) {
        job_t *j = &jobs[t];
        j->cfg = cfg; j->st = st; j->obs = obs;
        j->e0 = n_envs * t / n_threads; j->e1 = n_envs * (t + 1) / n_threads;
        j->env_id0 = env_id0; j->seed = seed; j->policy = policy; j->actions = actions;
        j->n_envs = n_envs; j->n_action_sets = n_action_sets > 0 ? n_action_sets : 1;
        j->scratch = actions_scratch; j->rewards = rewards; j->dones = dones;
        j->n_steps = n_steps; j->auto_reset = auto_reset;
    }
    for (int t = 1; t < n_threads; ++t) pthread_create(&tid[t], NULL, rollout_thread, &jobs[t]);
    rollout_thread(&jobs[0]);
    for (int t = 1; t < n_threads; ++t) pthread_join(tid[t], NULL);
    int64_t total = 0;
    for (int t = 0; t < n_threads; ++t) {
        total += jobs[t].agent_steps;
        if (stats) for (int i = 0; i < WHO_NUM_STATS; ++i) stats[i] += jobs[t].stats[i];
    }
    free(jobs); free(tid);
    return total;
}
