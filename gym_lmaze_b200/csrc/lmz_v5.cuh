// lmz_v5.cuh -- lmaze-v5 / lmaze-v6: the two-level planner / actor env
// (reference gym_lmaze/envs/lmaze_env_v5.py; v6 = v5 + safeFovealGoal(), lmaze_env_v6.py:505-523).
//
// Protocol:  reset() -> foveal obs (7,35,35)                       lmaze_env_v5.py:102-153
//                           plannerStep(goal25) -> local obs (4,35,35)            :158-182
//                           step(action4) -> foveal obs, local obs, globalReward, originalReward,
//                                            globalDone, localDone, fovealGoal, action   :187-292
// What shapes the kernel:
//   * the five 18x18 mazes of v2/v4 and v4's float visit layer state[2], but the layer is averaged with
//     the window around the ball only while the LOCAL episode is over (:308-312), and reset() zeroes it;
//   * two observation tensors per step: the foveal one has v4's structure (crop at the ball of free /
//     goal / visit, the fovealGoal one-hot, retStatelast = crop of the LIVE state where the last local
//     episode ended, :314-333,351-352); the local one is four 5x5 bit planes (free crop, ball and previous
//     ball relative to the planner-time fovea WITH numpy's negative-index wrap-around, fovealGoal,
//     :356-380).  Both are x7 upsampled and written as flat runs of float4 per 32-env tile, like v2/v4;
//   * ~80 bits of per-env state in three packed words.
// Where the reference raises IndexError (ball 3 cells right of / below the planner-time fovea, :365-366)
// the local obs row is written as all zeros, loc_err[e] = 1 and the error counter is bumped.
#pragma once
#include "lmz_v2.cuh"

namespace lmz {

enum : int { MODE_PLANNER = 3 };

struct V5 {
  static constexpr int ID = 5;
  static constexpr int G = 18, E = 7, F = 5, C = 7, CL = 4, S = F * E;       // lmaze_env_v5.py:25-37,357-358
  static constexpr int NLAYOUT = 5, MAX_CAND = 80;
  static constexpr int NBIT = 7;      // bit planes: free, goal, fovealGoal, last free, last goal, ball rel, previous ball rel
  static constexpr int NVIS = 2;      // float planes: visit crop at the ball, visit crop at retStatelast's window
  static constexpr bool MAZE_FIRST = true;                                   // reset(): setGrid() first (:104,115-116)
  static constexpr uint32_t OBS_FLOATS = C * S * S;                          // 8,575 (foveal)
  static constexpr uint32_t OBS_BYTES = OBS_FLOATS * 4;                      // 34,300
  static constexpr uint32_t LOC_FLOATS = CL * S * S;                         // 4,900 (local)
  static constexpr uint32_t LOC_BYTES = LOC_FLOATS * 4;                      // 19,600
  static constexpr uint32_t TILE_F4 = 32 * OBS_FLOATS / 4;                   // float4 per 32-env tile
  static constexpr uint32_t LOC_TILE_F4 = 32 * LOC_FLOATS / 4;
  static constexpr int STEP_LIMIT = 10, FSTEP_LIMIT = 50;                    // :47-48
  static constexpr uint32_t STEP_SAT = 255;
  // blob layout (bytes)
  static constexpr uint32_t LUT_OFF = 0;                                     // u8 [8575]: (slot << 5) | cell, foveal obs
  static constexpr uint32_t LOCLUT_OFF = align16(OBS_FLOATS);                // u8 [4900]: (bit plane << 5) | cell, local obs
  static constexpr uint32_t ROWBITS_OFF = LOCLUT_OFF + align16(LOC_FLOATS);  // u32 [5][18]
  static constexpr uint32_t CLS_OFF = ROWBITS_OFF + align16(NLAYOUT * G * 4);
  static constexpr uint32_t GCAND_OFF = CLS_OFF + align16(NLAYOUT * G * G);
  static constexpr uint32_t BCAND_OFF = GCAND_OFF + NLAYOUT * MAX_CAND * 2;
  static constexpr uint32_t BRANK_OFF = BCAND_OFF + NLAYOUT * MAX_CAND * 2;
  static constexpr uint32_t COUNT_OFF = BRANK_OFF + align16(NLAYOUT * G * G); // u8 ng[5], nb[5]; u16 xcell[5] at +16
  static constexpr uint32_t XCELL_OFF = COUNT_OFF + 16;
  static constexpr uint32_t BLOB_BYTES = COUNT_OFF + 32;
};

struct V5Regs {
  int L, x, y, gx, gy;        // maze, ball (= fovea_x0/y0), global goal
  int x1, y1;                 // previous ball            (lmaze_env_v5.py:192-193)
  int fx1, fy1;               // planner-time fovea       (:176-178)
  int fgx, fgy;               // f_goal_x0 / f_goal_y0    (:170-171)
  int lx, ly;                 // where retStatelast was taken (:327-328,351-352)
  int fga;                    // hot cell of fovealGoal   (:166-168; 12 after reset, :131-132)
  uint32_t step, fstep;       // stepCount, fovealStepCount (saturating; only >= 10 / >= 50 / == 0 matter)
  int ld, gd;                 // localDone, globalDone
};

// w0: L:3 | x:5 | y:5 | gx:5 | gy:5 | ld:1 | gd:1      w1: x1:5 | y1:5 | fx1:5 | fy1:5 | fgx:5 | fgy:5
// w2: lx:5 | ly:5 | fga:5 | step:8 | fstep:8
__host__ __device__ inline V5Regs v5_unpack(uint32_t w0, uint32_t w1, uint32_t w2) {
  V5Regs r;
  r.L = w0 & 7; r.x = (w0 >> 3) & 31; r.y = (w0 >> 8) & 31; r.gx = (w0 >> 13) & 31; r.gy = (w0 >> 18) & 31;
  r.ld = (w0 >> 23) & 1; r.gd = (w0 >> 24) & 1;
  r.x1 = w1 & 31; r.y1 = (w1 >> 5) & 31; r.fx1 = (w1 >> 10) & 31; r.fy1 = (w1 >> 15) & 31;
  r.fgx = (w1 >> 20) & 31; r.fgy = (w1 >> 25) & 31;
  r.lx = w2 & 31; r.ly = (w2 >> 5) & 31; r.fga = (w2 >> 10) & 31; r.step = (w2 >> 15) & 255; r.fstep = (w2 >> 23) & 255;
  return r;
}
__host__ __device__ inline void v5_pack(const V5Regs &r, uint32_t &w0, uint32_t &w1, uint32_t &w2) {
  w0 = (uint32_t)r.L | ((uint32_t)r.x << 3) | ((uint32_t)r.y << 8) | ((uint32_t)r.gx << 13) | ((uint32_t)r.gy << 18) |
       ((uint32_t)r.ld << 23) | ((uint32_t)r.gd << 24);
  w1 = (uint32_t)r.x1 | ((uint32_t)r.y1 << 5) | ((uint32_t)r.fx1 << 10) | ((uint32_t)r.fy1 << 15) |
       ((uint32_t)r.fgx << 20) | ((uint32_t)r.fgy << 25);
  w2 = (uint32_t)r.lx | ((uint32_t)r.ly << 5) | ((uint32_t)r.fga << 10) | (r.step << 15) | (r.fstep << 23);
}

// reset() (lmaze_env_v5.py:102-153): new maze, goal, ball; every history field collapses onto the ball
template <class W>
__device__ __forceinline__ void v5_respawn(V5Regs &r, const KParams &p, int64_t e, uint32_t &episode,
                                           const FovTables<W> &t, const unsigned char *smem) {
  V2Regs v;
  v.L = r.L; v.x = r.x; v.y = r.y; v.gx = r.gx; v.gy = r.gy; v.px = v.py = 0; v.a = -1; v.step = 0;
  v2_respawn<W>(v, p, e, episode, t);                     // MAZE_FIRST: maze, then goal, then ball (as v4)
  r.L = v.L; r.x = v.x; r.y = v.y; r.gx = v.gx; r.gy = v.gy;
  if (!p.random_goal) {                                   // setGoal() else-branch: the maze's 'X' cell (:483-485)
    const int xc = reinterpret_cast<const uint16_t *>(smem + W::XCELL_OFF)[r.L - 1];
    r.gx = xc / W::G; r.gy = xc % W::G;
  }
  if (!p.random_ball) { r.x = 4; r.y = 4; }               // setBall() else-branch: the 'S' cell, (4,4) in all five mazes
  r.x1 = r.fx1 = r.fgx = r.lx = r.x;                      // :139-148, and retStatelast = observation (:327-328)
  r.y1 = r.fy1 = r.fgy = r.ly = r.y;
  r.fga = 12;                                             // fovealGoal[0,2,2] = 1 (:131-132)
  r.step = 0; r.fstep = 0; r.ld = 0; r.gd = 0;            // :109-113
}

// numpy index of a 5-long axis: 0..4 direct, -5..-1 wrap, anything else raises IndexError (:365-366)
__device__ __forceinline__ bool v5_np_index(int &i) {
  if (i < -5 || i > 4) return false;
  if (i < 0) i += 5;
  return true;
}

struct V5Lane {
  LaneOut o;
  uint32_t mask[V5::NBIT];
  uint32_t info;            // x:5 | y:5 | shown last x:5 | shown last y:5 | visit op:2 (0 read, 1 average, 2 zero) | loc_err:1
  bool rfov, rloc;          // which observation rows this call writes
};

template <class W>
__device__ __forceinline__ V5Lane v5_lane(const KParams &p, int64_t e, const FovTables<W> &t, const unsigned char *smem) {
  V5Lane out;
  LaneOut &o = out.o;
  o.st_old = 0; o.render = false; o.done = false; o.cls = -1; o.eplen = 0;
  out.rfov = false; out.rloc = false;
  V5Regs r = v5_unpack(p.state[e], p.goal_count[e], p.aux2[e]);
  bool reset_now = false, write_state = false;
  uint32_t visit_op = 0;
  int slx = r.lx, sly = r.ly;                       // the retStatelast window the FOVEAL obs of this call shows
  if (p.mode == MODE_STEP) {
    const long long a = load_action(p.actions, p.action_dtype, e);
    r.x1 = r.x; r.y1 = r.y;                                                          // :192-193
    r.step = r.step < W::STEP_SAT ? r.step + 1 : W::STEP_SAT;                        // :196
    int dx = 0, dy = 0;                                                              // :203-217
    if (a == 0) dx = 1; else if (a == 1) dx = -1; else if (a == 2) dy = 1; else if (a == 3) dy = -1;
    const int nx = r.x + dx, ny = r.y + dy;
    const int tc = t.cls[(r.L - 1) * W::G * W::G + nx * W::G + ny];
    const bool at_fgoal = (nx == r.fgx && ny == r.fgy);
    const int was_gd = r.gd;
    int lcode = RC_NEG_ZERO, gcode;                                                  // :195
    if (tc == CLS_W) { lcode = RC_WALL; o.cls = CLS_W; }                             // :223-224
    else if (at_fgoal) { lcode = RC_GOAL; r.x = nx; r.y = ny; r.ld = 1; o.cls = CLS_B; }   // :226-230
    else {                                                                           // :232-242 (B, S or X)
      if (nx < r.fx1 - 3 || nx > r.fx1 + 2 || ny < r.fy1 - 3 || ny > r.fy1 + 2) r.ld = 1;
      lcode = RC_MOVE; r.x = nx; r.y = ny; o.cls = CLS_B;
    }
    if (nx == r.gx && ny == r.gy) { gcode = RC_GOAL; r.gd = 1; o.cls = CLS_X; }      // :245-247
    else if (at_fgoal) gcode = RC_MOVE;                                              // :248-249
    else gcode = RC_WALL;                                                            // :250-251
    if (r.step >= (uint32_t)W::STEP_LIMIT) r.ld = 1;                                 // :257-258
    if (r.fstep >= (uint32_t)W::FSTEP_LIMIT) { r.gd = 1; r.ld = 1; }                 // :260-262
    visit_op = r.ld ? 1u : 0u;                                                       // buildFovealObservation, :308-312
    if (r.fstep == 0) { r.lx = r.x; r.ly = r.y; }                                    // :327-328
    slx = r.lx; sly = r.ly;
    if (r.ld) { r.lx = r.x; r.ly = r.y; }                                            // :351-352, after the obs was built
    p.reward[e] = __uint_as_float(reward_bits(gcode));
    p.reward2[e] = __uint_as_float(reward_bits(lcode));
    p.done[e] = (uint8_t)r.gd;
    p.done2[e] = (uint8_t)r.ld;
    o.done = r.gd && !was_gd;
    if (o.done) o.eplen = r.fstep;
    reset_now = r.gd && p.autoreset;
    out.rfov = out.rloc = true;
    write_state = true;
  } else if (p.mode == MODE_PLANNER) {
    if (p.mask == nullptr || p.mask[e] != 0) {
      long long g = load_action(p.actions, p.action_dtype, e);
      if (g < 0 || g > 24) { atomicAdd(p.errors, 1u); g = g < 0 ? 0 : 24; }         // the reference raises IndexError (:168)
      r.step = 0; r.ld = 0;                                                          // :161-163
      r.fga = (int)g;                                                                // :165-168
      r.fgx = r.x + (int)g / 5 - 2; r.fgy = r.y + (int)g % 5 - 2;                    // :170-171
      if (r.fstep > 0) { r.fx1 = r.x; r.fy1 = r.y; }                                 // :176-178 (fovea_x0 is the ball)
      r.fstep = r.fstep < W::STEP_SAT ? r.fstep + 1 : W::STEP_SAT;                   // :180
      out.rloc = true;
      write_state = true;
    }
  } else if (p.mode == MODE_RESET) {
    reset_now = (p.mask == nullptr) || (p.mask[e] != 0);
    out.rfov = reset_now;
  } else {
    out.rfov = out.rloc = true;
  }
  if (reset_now) {
    uint32_t ep = p.episode[e];
    v5_respawn<W>(r, p, e, ep, t, smem);
    p.episode[e] = ep;
    visit_op = 2;                                   // self.state = zeros (:134); no averaging: localDone is False
    slx = r.lx; sly = r.ly;
    write_state = true;
  }
  // local observation (:356-380)
  int i0 = r.x - r.fx1 + 2, i1 = r.y - r.fy1 + 2, j0 = r.x1 - r.fx1 + 2, j1 = r.y1 - r.fy1 + 2;
  const bool loc_ok = v5_np_index(i0) & v5_np_index(i1) & v5_np_index(j0) & v5_np_index(j1);
  if (out.rloc) {
    if (!loc_ok) atomicAdd(p.errors, 1u);
    if (p.loc_err) p.loc_err[e] = loc_ok ? 0 : 1;
  }
  if (write_state) {
    uint32_t w0, w1, w2;
    v5_pack(r, w0, w1, w2);
    p.state[e] = w0; p.goal_count[e] = w1; p.aux2[e] = w2;
    if (p.fgoal_out) p.fgoal_out[e] = (uint8_t)r.fga;
  }
  o.st = 0;
  out.info = (uint32_t)r.x | ((uint32_t)r.y << 5) | ((uint32_t)slx << 10) | ((uint32_t)sly << 15) | (visit_op << 20) |
             (loc_ok ? 0u : (1u << 22));
  out.rfov = out.rfov && p.obs != nullptr;
  out.rloc = out.rloc && p.obs2 != nullptr;
  out.mask[0] = v2_free_crop<W>(t, r.L, r.x, r.y);                                   // state[0] crop at the fovea
  out.mask[1] = v2_goal_crop(r.x, r.y, r.gx, r.gy);                                  // state[1] crop
  out.mask[2] = 1u << r.fga;                                                         // fovealGoal
  out.mask[3] = v2_free_crop<W>(t, r.L, slx, sly);                                   // retStatelast (a view of state)
  out.mask[4] = v2_goal_crop(slx, sly, r.gx, r.gy);
  out.mask[5] = loc_ok ? (1u << (i0 * 5 + i1)) : 0u;                                 // :365
  out.mask[6] = loc_ok ? (1u << (j0 * 5 + j1)) : 0u;                                 // :366
  return out;
}

// Fused reset / plannerStep / step / render of the planner-actor env.  Same CTA organisation as the v4
// kernel: warps 0-3 are producers (warp 0 runs the next tile's transitions, then all four make the pass
// over that tile's float visit layers and capture the two 5x5 crops per env), the other warps write the
// current tile's foveal rows and then its local rows as flat float4 runs.
template <class W, int THREADS>
__global__ void __launch_bounds__(THREADS) lmz_env_v5_kernel(const KParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t s_mask[2][32 * W::NBIT];
  __shared__ uint32_t s_info[2][32];
  __shared__ float s_vis[2 * 32 * 50];       // [buf][env][0: at the ball, 1: at the shown retStatelast window][25]
  __shared__ uint32_t s_flags[2][2];         // [buf][0 foveal, 1 local]: bit l = env l of the tile is written
  __shared__ long long s_tile[2];
  constexpr int PROD = THREADS >= 512 ? 128 : 64;
  constexpr int CTHREADS = THREADS - PROD;
  static_assert(CTHREADS >= 32, "need at least one rendering warp");
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  stage_blob<W>(smem, &bar, p.blob);
  const FovTables<W> t(smem);
  const uint8_t *loclut = smem + W::LOCLUT_OFF;
  const int64_t tiles = p.tile_end;
  const bool need_visit = (p.mode != MODE_PLANNER);
  WarpStats ws;

  auto produce = [&](int buf) {              // warp 0 only
    int64_t tl = 0;
    if (lane == 0) tl = p.tile_begin + grab_tile(p.work);
    tl = __shfl_sync(0xffffffffu, tl, 0);
    const int64_t e = tl * 32 + lane;
    const bool valid = tl < tiles && e < p.n;
    V5Lane v;
    v.o.st = 0; v.o.st_old = 0; v.o.render = false; v.o.done = false; v.o.cls = -1; v.o.eplen = 0; v.info = 0;
    v.rfov = false; v.rloc = false;
#pragma unroll
    for (int c = 0; c < W::NBIT; ++c) v.mask[c] = 0;
    if (valid) v = v5_lane<W>(p, e, t, smem);
    if (p.mode == MODE_STEP) ws.add(valid, v.o);
#pragma unroll
    for (int c = 0; c < W::NBIT; ++c) s_mask[buf][lane * W::NBIT + c] = v.mask[c];
    s_info[buf][lane] = valid ? v.info : 0u;
    const unsigned ff = __ballot_sync(0xffffffffu, valid && v.rfov);
    const unsigned fl = __ballot_sync(0xffffffffu, valid && v.rloc);
    if (lane == 0) { s_flags[buf][0] = ff; s_flags[buf][1] = fl; s_tile[buf] = tl; }
  };
  // visit layers of one tile's 32 envs: 32 x 324 consecutive floats, coalesced pass by the producer threads
  auto visit_pass = [&](int buf) {
    const int64_t tile = s_tile[buf];
    if (tile >= tiles || !need_visit) return;
    const int64_t e0 = tile * 32;
    const uint32_t cells = (uint32_t)(((p.n - e0) < 32 ? (p.n - e0) : 32) * (W::G * W::G));
    float *vis = p.visit + e0 * (W::G * W::G);
    float *sv = s_vis + buf * (32 * 50);
    constexpr uint32_t PER = (32 * W::G * W::G + PROD - 1) / PROD, UN = 8;
    for (uint32_t k0 = 0; k0 < PER; k0 += UN) {
      float vv[UN];
#pragma unroll
      for (uint32_t j = 0; j < UN; ++j) {
        const uint32_t idx = tid + (k0 + j) * PROD;
        vv[j] = (idx < cells) ? __ldcs(vis + idx) : 0.0f;
      }
#pragma unroll
      for (uint32_t j = 0; j < UN; ++j) {
        const uint32_t idx = tid + (k0 + j) * PROD;
        if (idx >= cells) continue;
        const uint32_t env = idx / (W::G * W::G), cell = idx - env * (W::G * W::G);
        const int x = cell / W::G, y = cell - x * W::G;
        const uint32_t info = s_info[buf][env];
        const int bx = info & 31, by = (info >> 5) & 31, px = (info >> 10) & 31, py = (info >> 15) & 31;
        const uint32_t op = (info >> 20) & 3u;
        const int dx = x - bx + 2, dy = y - by + 2, qx = x - px + 2, qy = y - py + 2;
        const bool in_cur = dx >= 0 && dx < 5 && dy >= 0 && dy < 5;
        float v = vv[j];
        if (op == 1) v = (float)(((double)v + (in_cur ? 1.0 : 0.0)) * 0.5);       // :308-312, float64 then float32
        else if (op == 2) v = 0.0f;                                               // :134
        if (op) __stcs(vis + idx, v);
        if (in_cur) sv[env * 50 + dx * 5 + dy] = v;
        if (qx >= 0 && qx < 5 && qy >= 0 && qy < 5) sv[env * 50 + 25 + qx * 5 + qy] = v;
      }
    }
  };
  auto producers_sync = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(PROD) : "memory"); };

  if (warp == 0) produce(0);
  if (tid < PROD) { producers_sync(); visit_pass(0); }
  for (int buf = 0;; buf ^= 1) {
    __syncthreads();                          // tile(buf) is complete; buffers buf^1 are free again
    const int64_t tile = s_tile[buf];
    if (tile >= tiles) break;
    if (tid < PROD) {
      if (warp == 0) produce(buf ^ 1);
      producers_sync();
      visit_pass(buf ^ 1);
      continue;
    }
    const int ctid = tid - PROD;
    const uint32_t *mk = s_mask[buf];
    const uint32_t *inf = s_info[buf];
    const float *sv = s_vis + buf * (32 * 50);
    const int64_t row0 = tile * 32 - p.win_lo;
    // ---- foveal rows (:314-348)
    uint32_t flags = s_flags[buf][0];
    if (flags) {
      float *dst = reinterpret_cast<float *>(p.obs) + row0 * (int64_t)W::OBS_FLOATS;
      auto value = [&](uint32_t env, uint32_t r) -> uint32_t {
        const uint32_t code = t.lut[r], slot = code >> 5, cell = code & 31u;
        if (slot >= 5u) return __float_as_uint(sv[env * 50 + (slot - 5u) * 25 + cell]);
        return ((mk[env * W::NBIT + slot] >> cell) & 1u) ? 0x3f800000u : 0u;
      };
      if (flags == 0xffffffffu && (row0 & 3) == 0) {
        for (uint32_t q = ctid; q < W::TILE_F4; q += CTHREADS) {
          uint32_t g = q * 4, env = g / W::OBS_FLOATS, r = g - env * W::OBS_FLOATS;
          uint4 v;
          uint32_t *w = reinterpret_cast<uint32_t *>(&v);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            w[k] = value(env, r);
            if (++r == W::OBS_FLOATS) { r = 0; ++env; }
          }
          st_stream_v4(reinterpret_cast<unsigned char *>(dst) + ((size_t)q << 4), v);
        }
      } else {
        for (uint32_t g = ctid; g < 32 * W::OBS_FLOATS; g += CTHREADS) {
          const uint32_t env = g / W::OBS_FLOATS, r = g - env * W::OBS_FLOATS;
          if ((flags >> env) & 1u) __stcs(reinterpret_cast<unsigned int *>(dst) + g, value(env, r));
        }
      }
    }
    // ---- local rows (:356-380): every env row is 16-byte aligned (19,600 = 16 x 1,225)
    flags = s_flags[buf][1];
    if (flags) {
      float *dst = reinterpret_cast<float *>(p.obs2) + row0 * (int64_t)W::LOC_FLOATS;
      auto value = [&](uint32_t env, uint32_t r) -> uint32_t {
        const uint32_t code = loclut[r], plane = code >> 5, cell = code & 31u;
        return ((mk[env * W::NBIT + plane] >> cell) & 1u) ? 0x3f800000u : 0u;
      };
      for (uint32_t q = ctid; q < W::LOC_TILE_F4; q += CTHREADS) {
        const uint32_t env = q / (W::LOC_FLOATS / 4), r = (q - env * (W::LOC_FLOATS / 4)) * 4;
        if (!((flags >> env) & 1u)) continue;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (!((inf[env] >> 22) & 1u)) {          // IndexError in the reference: the row stays all zero
          v.x = value(env, r); v.y = value(env, r + 1); v.z = value(env, r + 2); v.w = value(env, r + 3);
        }
        st_stream_v4(reinterpret_cast<unsigned char *>(dst) + ((size_t)q << 4), v);
      }
    }
  }
  if (warp == 0) {
    if (p.mode == MODE_STEP) ws.flush(p.stats, lane);
    if (lane == 0) finish_grabber(p.work, (unsigned long long)gridDim.x);
  }
}

// v6 safeFovealGoal() (lmaze_env_v6.py:505-523): a foveal goal whose window cell is not a wall.
// draws != null: int64 [n][n_draws] values the reference's np.random.randint(0, 25) would have returned;
// the first non-wall one wins (used[e] = how many were consumed; goal -1 if none qualified).
// draws == null: exactly uniform over the window's non-wall cells (the distribution the rejection loop
// has), Philox keyed by (seed, global env id, episode, fovealStepCount).
constexpr uint32_t TAG_SAFE = 0x47u;
template <class W>
__global__ void lmz_safe_goal_kernel(int64_t n, const uint32_t *state, const uint32_t *aux2, const uint32_t *episode,
                                     const uint8_t *blob, const long long *draws, int n_draws, uint8_t *goal_out,
                                     int32_t *used_out, uint64_t seed, uint64_t env_id0) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const uint32_t w0 = state[e];
  const int L = w0 & 7, x = (w0 >> 3) & 31, y = (w0 >> 8) & 31;
  const uint32_t *rowbits = reinterpret_cast<const uint32_t *>(blob + W::ROWBITS_OFF);
  uint32_t m = 0;                                               // smallGrid != 'W', bit i*5+j
#pragma unroll
  for (int i = 0; i < 5; ++i) m |= ((rowbits[(L - 1) * W::G + x - 2 + i] >> (y - 2)) & 31u) << (5 * i);
  int goal = -1, used = 0;
  if (draws) {
    for (int k = 0; k < n_draws; ++k) {
      const long long a = draws[e * n_draws + k];
      used = k + 1;
      if (a >= 0 && a < 25 && ((m >> a) & 1u)) { goal = (int)a; break; }
    }
  } else {
    const int cnt = __popc(m);                                  // >= 1: the ball's own cell is never a wall
    const uint32_t fstep = (aux2[e] >> 23) & 255u;
    const uint64_t gid = env_id0 + (uint64_t)e;
    uint32_t out[4];
    uint32_t k = 0;
    for (uint32_t attempt = 0;; ++attempt) {
      philox4x32_10((uint32_t)gid, (uint32_t)(gid >> 32), episode[e], (attempt << 16) | (fstep << 8) | TAG_SAFE,
                    (uint32_t)seed, (uint32_t)(seed >> 32), out);
      const uint32_t thresh = (uint32_t)(0x100000000ull % (uint32_t)cnt);
      bool got = false;
#pragma unroll
      for (int i = 0; i < 4 && !got; ++i) {
        const uint64_t mm = (uint64_t)out[i] * (uint64_t)cnt;
        if ((uint32_t)mm >= thresh) { k = (uint32_t)(mm >> 32); got = true; }
      }
      if (got) break;
    }
    uint32_t mm = m;
    for (uint32_t i = 0; i < k; ++i) mm &= mm - 1;              // k-th set bit
    goal = __ffs(mm) - 1;
  }
  goal_out[e] = (uint8_t)(goal < 0 ? 255 : goal);
  if (used_out) used_out[e] = used;
}

// get/set_state for v5/v6.  int32 [n][16]: x, y, x1, y1, fx1, fy1, gx, gy, fgx, fgy, last_x, last_y, fgoal_action,
// step, foveal_step, flags (globalDone | localDone << 1 | layout << 4);
// the episode counter travels in a 17th column.
__global__ void lmz_state_v5_kernel(int64_t n, uint32_t *state, uint32_t *aux1, uint32_t *aux2, uint32_t *episode,
                                    int32_t *io, int set) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  int32_t *row = io + e * 17;
  if (set) {
    auto pos = [](int v) { return v < 2 ? 2 : (v > V5::G - 3 ? V5::G - 3 : v); };     // 5x5 crops stay inside the grid
    V5Regs r;
    r.x = pos(row[0]); r.y = pos(row[1]); r.x1 = pos(row[2]); r.y1 = pos(row[3]); r.fx1 = pos(row[4]); r.fy1 = pos(row[5]);
    r.gx = row[6] & 31; r.gy = row[7] & 31; r.fgx = row[8] & 31; r.fgy = row[9] & 31; r.lx = pos(row[10]); r.ly = pos(row[11]);
    r.fga = row[12] < 0 ? 0 : (row[12] > 24 ? 24 : row[12]);
    r.step = (uint32_t)row[13] < V5::STEP_SAT ? (uint32_t)row[13] : V5::STEP_SAT;
    r.fstep = (uint32_t)row[14] < V5::STEP_SAT ? (uint32_t)row[14] : V5::STEP_SAT;
    r.gd = row[15] & 1; r.ld = (row[15] >> 1) & 1;
    const int L = (row[15] >> 4) & 7;
    r.L = L < 1 ? 1 : (L > 5 ? 5 : L);
    uint32_t w0, w1, w2;
    v5_pack(r, w0, w1, w2);
    state[e] = w0; aux1[e] = w1; aux2[e] = w2; episode[e] = (uint32_t)row[16];
  } else {
    const V5Regs r = v5_unpack(state[e], aux1[e], aux2[e]);
    row[0] = r.x; row[1] = r.y; row[2] = r.x1; row[3] = r.y1; row[4] = r.fx1; row[5] = r.fy1; row[6] = r.gx; row[7] = r.gy;
    row[8] = r.fgx; row[9] = r.fgy; row[10] = r.lx; row[11] = r.ly; row[12] = r.fga; row[13] = (int32_t)r.step;
    row[14] = (int32_t)r.fstep; row[15] = r.gd | (r.ld << 1) | (r.L << 4); row[16] = (int32_t)episode[e];
  }
}

}  // namespace lmz
