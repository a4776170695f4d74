/*
 * lmaze_oracle.h -- CPU restatement of gkm2708/gym-lmaze's step()/reset() hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library, and only as the
 * checker or the CPU baseline -- never as the thing shipped.  The product
 * (gym_lmaze_b200/) does not link, import or fall back to anything in oracle/.
 *
 * Parity pin: the reference ships NO tests, golden vectors or fixtures
 * (SURVEY.md section 8c).  This restatement is therefore pinned against outputs
 * of the unmodified reference itself, executed in the build container under a
 * `gym` stub: tests/golden/gen_golden.py wrote the fixtures in tests/golden/
 * (exhaustive v0 transition table, all position->obs renders, seeded traces,
 * v3 tables) and tests/test_oracle_golden.py checks every one of them; where
 * /root/reference is present tests/test_oracle_vs_reference.py also steps the
 * live reference beside this code.
 *
 * All file:line citations are into the reference checkout.
 */
#ifndef LMAZE_ORACLE_H
#define LMAZE_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LMZO_V0 0
#define LMZO_V2 2
#define LMZO_V3 3
#define LMZO_V4 4
#define LMZO_V5 5   /* also stands for v6, which only adds safeFovealGoal() */
#define LMZO_MAX_G 18

/* One environment, holding the same mutable fields the reference keeps on `self`. */
typedef struct lmzo_env {
  int variant;            /* LMZO_V0 / LMZO_V3 */
  int G;                  /* realgrid          (lmaze_env.py:17, lmaze_env_v3.py:72) */
  int E;                  /* expansionRatio    (lmaze_env.py:18, lmaze_env_v3.py:73) */
  int C;                  /* obs channels      (lmaze_env.py:20, lmaze_env_v3.py:87-89) */
  int step_limit;         /* 100 both variants (lmaze_env.py:247, lmaze_env_v3.py:99)  */
  int random_ball;        /* RANDOM_BALL       (lmaze_env.py:25, lmaze_env_v3.py:103)  */
  int random_goal;        /* RANDOM_GOAL, v3   (lmaze_env_v3.py:104) */
  char grid[LMZO_MAX_G * LMZO_MAX_G];         /* row-major cell letters */
  float state[4 * LMZO_MAX_G * LMZO_MAX_G];   /* self.state layers (v0: 4 layers; v3: 2 used) */
  float goal_image[LMZO_MAX_G * LMZO_MAX_G];  /* v3 image_global_goal (lmaze_env_v3.py:170,182) */
  int ball_x, ball_y;     /* ball_x0, ball_y0 (row, col) */
  int goal_x, goal_y;
  double reward;          /* Python float in the reference */
  int64_t step_count;     /* stepCount */
  int64_t goal_count;     /* goalCount, v0 only; survives reset (lmaze_env.py:24,195) */
  /* ---- v2 (lmaze_env_v2.py): multi-layout foveal env ---- */
  int layout;             /* 1..5: which maze setGrid() picked (lmaze_env_v2.py:303-405) */
  int fovea;              /* 5 (lmaze_env_v2.py:27) */
  float last_crop[2 * 25];/* retStatelast: the previous 2x5x5 observation (lmaze_env_v2.py:41,104,212-213) */
  float shown_prev[2 * 25];/* the retStatelast that went into the CURRENT obs (before :212-213 replaced it) */
  float action_plane[25]; /* action_value one-hot (lmaze_env_v2.py:87,136-137) */
  int prev_x, prev_y;     /* where shown_prev was taken (bookkeeping for tests) */
  int64_t bad_actions;    /* batched API: actions outside 0..24 are clamped and counted */
  /* ---- v5 / v6 (lmaze_env_v5.py): planner / actor protocol ---- */
  int ball_x1, ball_y1;   /* previous ball (lmaze_env_v5.py:192-193) */
  int fovea_x1, fovea_y1; /* fovea centre at the last plannerStep (:176-178); fovea_x0/y0 always equal the ball */
  int fgoal_x, fgoal_y;   /* f_goal_x0/y0: the cell the planner pointed at (:170-171) */
  int fgoal_action;       /* which cell of fovealGoal is hot (:166-168); 12 after reset (:131-132) */
  int last_x, last_y;     /* where retStatelast (a VIEW of state) was taken (:327-328,:351-352) */
  int64_t foveal_step_count;
  int local_done, global_done;
  double global_reward;
} lmzo_env;

/* Episode statistics, integers only (order-independent sums). */
enum {
  LMZO_STAT_STEPS = 0, LMZO_STAT_EPISODES, LMZO_STAT_GOALS, LMZO_STAT_TIMEOUTS,
  LMZO_STAT_WALL_BUMPS, LMZO_STAT_MOVES, LMZO_STAT_STALE, LMZO_STAT_EPLEN_SUM,
  LMZO_NUM_STATS
};

int  lmzo_abi_version(void);

/* Layout access (lmaze_env.py:37-48, lmaze_env_v3.py:26-43).  Returns G, fills cells[G*G]. */
int  lmzo_layout(int variant, char *cells);
/* Sizes: obs floats per env. */
int64_t lmzo_obs_floats(int variant);

/* Constructor state minus the first reset (lmaze_env.py:14-50, lmaze_env_v3.py:22-131). */
int  lmzo_init(lmzo_env *e, int variant);

/* reset() with the rejection-sampled coordinates supplied by the caller.
 * v0: (sx,sy) is the ball spawn; gx,gy ignored.            lmaze_env.py:64-111
 * v3: (gx,gy) goal then (sx,sy) ball; pass gx<0 to keep the
 *     current goal (RANDOM_GOAL False).                     lmaze_env_v3.py:134-182
 * Returns 0, or -1 if the reference's rejection loop would not have accepted
 * the supplied cell. */
int  lmzo_reset(lmzo_env *e, int sx, int sy, int gx, int gy);

/* step(): action codes 0..3 move, anything else is the reference's unmatched
 * branch (offset 0,0).  For v3 the code k stands for the string str(k)
 * (lmaze_env_v3.py:236-247).  Returns done (0/1); reward is left in e->reward.
 * `cls_out` (may be NULL) receives the branch taken: 0 W, 1 B/else, 2 X/goal, 3 none(S). */
int  lmzo_step(lmzo_env *e, int64_t action, int *cls_out);

/* v2 reset: goal (gx,gy) and ball (sx,sy) as drawn on the CURRENT maze, then the maze is
 * re-rolled to `new_layout` (1..5) -- the order of lmaze_env_v2.py:90-92.  -1 if the reference's
 * rejection loops (:277-299) would not have accepted the cells. */
int  lmzo_reset_v2(lmzo_env *e, int sx, int sy, int gx, int gy, int new_layout);
/* v4 reset (lmaze_env_v4.py:89-149): the maze is re-rolled FIRST, goal and ball are drawn on the
 * NEW maze, state[2] (the visit layer) restarts at visitMap/2. */
int  lmzo_reset_v4(lmzo_env *e, int sx, int sy, int gx, int gy, int new_layout);
void lmzo_rng_spawn_v4(uint64_t seed, uint64_t env_id, uint32_t episode,
                       int *sx, int *sy, int *gx, int *gy, int *new_layout);
/* v4 visit layer state[2] (float32 [G*G]) of every env: read / overwrite */
void lmzo_vec_export_visit(const lmzo_env *envs, int64_t n, float *out);
void lmzo_env_set_visit(lmzo_env *e, const float *visit);
/* ---- v5 / v6 ---- */
int  lmzo_reset_v5(lmzo_env *e, int sx, int sy, int gx, int gy, int new_layout);   /* lmaze_env_v5.py:102-153 */
int  lmzo_planner_v5(lmzo_env *e, int64_t goal);                                   /* plannerStep, :158-182 */
/* step(), :187-292.  Returns global_done | local_done << 1; rewards are left in e->reward (originalReward)
 * and e->global_reward.  Also performs the state change of the buildFovealObservation() call inside step
 * (the visit-layer update when the local episode is over, :308-312, and retStatelast, :351-352). */
int  lmzo_step_v5(lmzo_env *e, int64_t action);
void lmzo_render_fov_v5(const lmzo_env *e, float *obs);      /* (7,35,35), :306-354 */
int  lmzo_render_loc_v5(const lmzo_env *e, float *obs);      /* (4,35,35), :356-380; -1 where the reference raises IndexError */
int  lmzo_step_full_v5(lmzo_env *e, int64_t action, float *fov, float *loc, int *loc_error);
/* Batched drivers (serial).  mask NULL = all envs.  spawn: int32 [N][4] (sx, sy, gx, gy | layout << 5) or NULL =
 * the Philox spec (same stream as v4: maze, goal, ball).  loc_err[i] = 1 where the reference would raise. */
void lmzo_vec_reset_v5(lmzo_env *envs, int64_t n, const uint8_t *mask, const int32_t *spawn, uint64_t seed,
                       uint64_t env_id0, uint32_t *episode, float *fov);
void lmzo_vec_planner_v5(lmzo_env *envs, int64_t n, const int64_t *goals, const uint8_t *mask, float *loc,
                         uint8_t *loc_err);
void lmzo_vec_step_v5(lmzo_env *envs, int64_t n, const int64_t *actions, const uint8_t *mask, float *fov, float *loc,
                      float *greward, float *lreward, uint8_t *gdone, uint8_t *ldone, uint8_t *loc_err);
void lmzo_vec_render_v5(const lmzo_env *envs, int64_t n, const uint8_t *mask, float *fov, float *loc, uint8_t *loc_err);
/* int32 [N][16]: x, y, x1, y1, fx1, fy1, gx, gy, fgx, fgy, last_x, last_y, fgoal_action, step, foveal_step, flags */
void lmzo_vec_export_v5(const lmzo_env *envs, int64_t n, int32_t *out);
/* v6 safeFovealGoal (lmaze_env_v6.py:505-523): consumes draws[] (each in 0..24) until a non-wall cell of
 * the 5x5 window around the ball comes up; returns it and stores how many draws were used. */
int  lmzo_safe_goal_v6(const lmzo_env *e, const int64_t *draws, int n_draws, int *used);
/* The five 18x18 mazes of lmaze_env_v2.py:303-405 (layout 1..5). */
int  lmzo_layout_v2(int layout, char *cells);

/* isEpisodeFinished() / the done expression (lmaze_env.py:246-249, lmaze_env_v3.py:398). */
int  lmzo_done(const lmzo_env *e);

/* The observation the reference returns: channel assembly + nested xE upsample
 * (lmaze_env.py:113-139,208-234; lmaze_env_v3.py:184-206,281-301). obs = C*G*E*G*E floats. */
void lmzo_render(const lmzo_env *e, float *obs);

/* ---- Philox-4x32-10 counter RNG: NOT from the reference (which uses CPython's
 * MT19937, lmaze_env.py:74-75).  It restates the device RNG spec in DESIGN.md so
 * RNG-driven spawns/actions of the CUDA path can be replayed on the CPU. ---- */
void lmzo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* Spawn for (seed, global env id, episode index).  v0: ball only.  v3: goal + ball. */
void lmzo_rng_spawn(int variant, uint64_t seed, uint64_t env_id, uint32_t episode,
                    int *sx, int *sy, int *gx, int *gy);
void lmzo_rng_spawn_v2(uint64_t seed, uint64_t env_id, uint32_t episode, int cur_layout,
                       int *sx, int *sy, int *gx, int *gy, int *first_layout, int *new_layout);
/* Rollout action for (seed, global env id, global rollout step t). */
int  lmzo_rng_action(uint64_t seed, uint64_t env_id, uint64_t t);
/* The same for the Discrete(25) variants (v2 / v4) and for the planner / actor env (goal 0..24 + action 0..3). */
int  lmzo_rng_action25(uint64_t seed, uint64_t env_id, uint64_t t);
void lmzo_rng_hier(uint64_t seed, uint64_t env_id, uint64_t t, int *goal, int *action);

/* ---- Vectorised driver: the batched semantics of the framework composed from
 * the single-env restatement above.  For each env i: step; record f32 reward and
 * done; if done and autoreset, reset (spawn from `spawn` if non-NULL, else the
 * Philox spec, bumping episode[i]); render into obs (if non-NULL).
 * spawn layout: int32 [N][4] = sx, sy, gx, gy (v2: column 3 = gy | new_layout << 5).
 * actions: int64 [N].  stats: int64[LMZO_NUM_STATS], accumulated.  threads<=1 => serial; else a pthread pool over contiguous env ranges. */
void lmzo_vec_step(lmzo_env *envs, int64_t n, const int64_t *actions, const int32_t *spawn,
                   uint64_t seed, uint64_t env_id0, uint32_t *episode, int autoreset,
                   float *obs, float *reward, uint8_t *done, int64_t *stats, int threads);
void lmzo_vec_reset(lmzo_env *envs, int64_t n, int variant, const int32_t *spawn,
                    uint64_t seed, uint64_t env_id0, uint32_t *episode, float *obs, int threads);

/* ---- array plumbing used by the ctypes wrapper (oracle/oracle.py) ---- */
int64_t lmzo_sizeof_env(void);
lmzo_env *lmzo_env_at(lmzo_env *envs, int64_t i);
void lmzo_env_set_flags(lmzo_env *e, int random_ball, int random_goal);
/* pos: int32 [N][4] = x, y, goal_x, goal_y */
void lmzo_vec_export(const lmzo_env *envs, int64_t n, int32_t *pos, int64_t *step_count,
                     int64_t *goal_count, double *reward);
/* aux: int32 [N][4] = layout, prev_x, prev_y, bad_actions (v2) */
void lmzo_vec_export_aux(const lmzo_env *envs, int64_t n, int32_t *aux);
/* v2: force maze, ball, goal, previous-crop position and stepCount */
int  lmzo_env_force_v2(lmzo_env *e, int layout, int sx, int sy, int gx, int gy, int px, int py, int64_t step_count);
/* Put one env into an arbitrary reachable mid-episode state (incl. ball on the goal cell). */
int  lmzo_env_force(lmzo_env *e, int variant, int sx, int sy, int gx, int gy,
                    int64_t step_count, double reward, int64_t goal_count);

#ifdef __cplusplus
}
#endif
#endif
