// lmz_v2.cuh -- lmaze-v2: multi-layout foveal env (reference gym_lmaze/envs/lmaze_env_v2.py).
//
// Differences from v0/v3 that shape the kernel:
//   * five 18x18 mazes, one re-rolled per episode (lmaze_env_v2.py:92,303-405): the per-layout
//     tables (row bitmasks, cell classes, goal / ball candidate lists) are staged into shared
//     memory with the same single bulk async (TMA) load as the other variants' blob;
//   * Discrete(25) "teleport the fovea" actions (:39,151-169), reward by the cell the fovea was
//     pointed at (:175-180), done at reward 100 or stepCount > 50 (:222);
//   * the observation is five 5x5 BIT planes -- free-cell crop and goal crop around the ball, the
//     action one-hot, and the previous step's two crops (:185-193) -- each upsampled x7 to 35x35
//     floats: 6,125 floats = 24,500 B per env.  24,500 is not a multiple of 16, so neither bulk
//     copies nor per-env vector stores are possible; instead a TILE of 32 envs (784,000 B, which
//     IS 16-byte aligned) is written as one flat run of float4 stores, every float looked up as
//     bit `cell` of mask `c` of env `e` through a 6,125-entry (c,cell) table in shared memory.
#pragma once
#include "lmz_kernels.cuh"

namespace lmz {

// Shared blob layout of the foveal variants (v2, v4); only the LUT length differs.
template <int ID_, int C_>
struct Fov {
  static constexpr int ID = ID_;
  static constexpr int G = 18, E = 7, F = 5, C = C_, S = F * E;       // lmaze_env_v2.py:26-37, lmaze_env_v4.py:26-40
  static constexpr int NLAYOUT = 5, MAX_CAND = 80;
  static constexpr int NBIT = 5;                                       // bit planes: free, goal, action, prev free, prev goal
  static constexpr int NVIS = (C_ == 7) ? 2 : 0;                       // v4: float visit crop at the ball and at the previous window
  static constexpr bool MAZE_FIRST = (ID_ == 4);                       // v4 re-rolls the maze BEFORE drawing goal/ball (:91-98)
  static constexpr uint32_t OBS_FLOATS = C * S * S;                    // 6,125 (v2) / 8,575 (v4)
  static constexpr uint32_t OBS_BYTES = OBS_FLOATS * 4;                // 24,500 / 34,300
  static constexpr uint32_t TILE_F4 = 32 * OBS_FLOATS / 4;             // float4 per 32-env tile (tile is 16-B aligned)
  static constexpr int STEP_LIMIT = 50;                                // lmaze_env_v2.py:43, lmaze_env_v4.py:48
  static constexpr uint32_t STEP_SAT = 63;
  // blob layout (bytes)
  static constexpr uint32_t LUT_OFF = 0;                               // u8 [OBS_FLOATS]: (slot << 5) | cell
  static constexpr uint32_t ROWBITS_OFF = align16(OBS_FLOATS);         // u32 [5][18]: bit y = cell (x,y) is B/S/X
  static constexpr uint32_t CLS_OFF = ROWBITS_OFF + align16(NLAYOUT * G * 4);     // u8 [5][324]
  static constexpr uint32_t GCAND_OFF = CLS_OFF + align16(NLAYOUT * G * G);       // u16 [5][80] cells not in {W,S}
  static constexpr uint32_t BCAND_OFF = GCAND_OFF + NLAYOUT * MAX_CAND * 2;       // u16 [5][80] cells not in {W,X}
  static constexpr uint32_t BRANK_OFF = BCAND_OFF + NLAYOUT * MAX_CAND * 2;       // i8 [5][324] index in BCAND or -1
  static constexpr uint32_t COUNT_OFF = BRANK_OFF + align16(NLAYOUT * G * G);     // u8 ng[5], nb[5]
  static constexpr uint32_t BLOB_BYTES = COUNT_OFF + 32;
};
using V2 = Fov<2, 5>;
using V4 = Fov<4, 7>;

struct V2Regs {
  int L, x, y, gx, gy, px, py, a;     // a: last action, -1 right after a reset (action plane all zero)
  uint32_t step;
};

// state word: L:3 | x:5 | y:5 | gx:5 | gy:5 | step:6 ; aux word: px:5 | py:5 | a:5 | a_valid:1
__host__ __device__ inline V2Regs v2_unpack(uint32_t s, uint32_t aux) {
  V2Regs r;
  r.L = s & 7; r.x = (s >> 3) & 31; r.y = (s >> 8) & 31; r.gx = (s >> 13) & 31; r.gy = (s >> 18) & 31;
  r.step = (s >> 23) & 63;
  r.px = aux & 31; r.py = (aux >> 5) & 31; r.a = ((aux >> 15) & 1) ? (int)((aux >> 10) & 31) : -1;
  return r;
}
__host__ __device__ inline void v2_pack(const V2Regs &r, uint32_t &s, uint32_t &aux) {
  s = (uint32_t)r.L | ((uint32_t)r.x << 3) | ((uint32_t)r.y << 8) | ((uint32_t)r.gx << 13) | ((uint32_t)r.gy << 18) |
      (r.step << 23);
  aux = (uint32_t)r.px | ((uint32_t)r.py << 5) | (r.a >= 0 ? (((uint32_t)r.a << 10) | (1u << 15)) : 0u);
}

template <class W>
struct FovTables {
  const uint8_t *lut;
  const uint32_t *rowbits;
  const uint8_t *cls;
  const uint16_t *gcand, *bcand;
  const int8_t *brank;
  const uint8_t *count;      // ng[0..4], nb[5..9]
  __device__ __forceinline__ explicit FovTables(const unsigned char *smem)
      : lut(smem + W::LUT_OFF), rowbits(reinterpret_cast<const uint32_t *>(smem + W::ROWBITS_OFF)),
        cls(smem + W::CLS_OFF), gcand(reinterpret_cast<const uint16_t *>(smem + W::GCAND_OFF)),
        bcand(reinterpret_cast<const uint16_t *>(smem + W::BCAND_OFF)),
        brank(reinterpret_cast<const int8_t *>(smem + W::BRANK_OFF)), count(smem + W::COUNT_OFF) {}
};

// 5x5 crop of the free-cell layer around (x,y) as 25 bits, bit i*5+j = cell (x-2+i, y-2+j)
template <class W>
__device__ __forceinline__ uint32_t v2_free_crop(const FovTables<W> &t, int L, int x, int y) {
  uint32_t m = 0;
#pragma unroll
  for (int i = 0; i < 5; ++i) m |= ((t.rowbits[(L - 1) * V2::G + x - 2 + i] >> (y - 2)) & 31u) << (5 * i);
  return m;
}
__device__ __forceinline__ uint32_t v2_goal_crop(int x, int y, int gx, int gy) {
  const int dx = gx - (x - 2), dy = gy - (y - 2);
  return (dx >= 0 && dx < 5 && dy >= 0 && dy < 5) ? (1u << (dx * 5 + dy)) : 0u;
}

// reset().  v2: goal and ball are drawn on the maze of the episode that just ended, THEN the maze is
// re-rolled (lmaze_env_v2.py:90-92,277-299).  v4: the maze is re-rolled first and goal/ball are drawn
// on the new one (lmaze_env_v4.py:91-98).  Injected spawn: sx, sy, gx, gy | new_layout << 5.
template <class W>
__device__ __forceinline__ void v2_respawn(V2Regs &r, const KParams &p, int64_t e, uint32_t &episode,
                                           const FovTables<W> &t) {
  int sx, sy, gx, gy, nl;
  if (p.spawn) {
    const int4 s = p.spawn[e];
    sx = s.x; sy = s.y; gx = s.z; gy = s.w & 31; nl = s.w >> 5;
    bool ok = nl >= 1 && nl <= 5;
    const int Lc = W::MAZE_FIRST ? (ok ? nl : 1) : r.L;           // the maze the rejection loops look at
    const uint8_t *cls = t.cls + (Lc - 1) * W::G * W::G;
    if (ok) ok = gx >= 1 && gx <= W::G - 2 && gy >= 1 && gy <= W::G - 2 && sx >= 1 && sx <= W::G - 2 && sy >= 1 &&
                 sy <= W::G - 2;
    if (ok) {
      const int gc = cls[gx * W::G + gy], bc = cls[sx * W::G + sy];
      ok = gc != CLS_W && gc != CLS_S && bc != CLS_W && bc != CLS_X && !(sx == gx && sy == gy);
    }
    if (!ok) {                                    // rejected: count it, fall back to the S / X cells of maze 1
      atomicAdd(p.errors, 1u);
      sx = 4; sy = 4; gx = 8; gy = 8; nl = 1;
    }
  } else {
    WordStream ws;
    ws.init(p.seed, p.env_id0 + (uint64_t)e, episode);
    int L0;
    if (W::MAZE_FIRST) { nl = 1 + (int)ws.uniform(5); L0 = nl - 1; }
    else {
      if (episode == 0) r.L = 1 + (int)ws.uniform(5);              // the maze the constructor rolled (:75)
      L0 = r.L - 1;
    }
    const int g = t.gcand[L0 * W::MAX_CAND + ws.uniform(t.count[L0])];
    const int rank = t.brank[L0 * W::G * W::G + g];
    const uint32_t nb = t.count[5 + L0];
    uint32_t k;
    if (rank >= 0) { k = ws.uniform(nb - 1); if ((int)k >= rank) k += 1; }
    else k = ws.uniform(nb);
    const int b = t.bcand[L0 * W::MAX_CAND + k];
    gx = g / W::G; gy = g % W::G; sx = b / W::G; sy = b % W::G;
    if (!W::MAZE_FIRST) nl = 1 + (int)ws.uniform(5);
  }
  r.L = nl; r.x = sx; r.y = sy; r.gx = gx; r.gy = gy; r.px = sx; r.py = sy; r.a = -1; r.step = 0;
  episode += 1;
}

struct V2Lane {
  LaneOut o;
  uint32_t mask[5];         // the five 25-bit planes: free, goal, action, previous free, previous goal
  uint32_t info;            // x:5 | y:5 | px:5 | py:5 | visit op:2  (v4: 0 keep, 1 step update, 2 reset)
};

template <class W>
__device__ __forceinline__ V2Lane v2_lane(const KParams &p, int64_t e, const FovTables<W> &t) {
  V2Lane out;
  LaneOut &o = out.o;
  o.st_old = 0; o.render = false; o.done = false; o.cls = -1; o.eplen = 0;
  V2Regs r = v2_unpack(p.state[e], p.goal_count[e]);
  bool reset_now = false;
  int reward_code = RC_NEG_ZERO;
  uint32_t visit_op = 0;
  if (p.mode == MODE_STEP) {
    visit_op = 1;
    long long a64 = load_action(p.actions, p.action_dtype, e);
    if (a64 < 0 || a64 > 24) { atomicAdd(p.errors, 1u); a64 = a64 < 0 ? 0 : 24; }   // reference raises IndexError
    const int a = (int)a64;
    r.step = r.step < W::STEP_SAT ? r.step + 1 : W::STEP_SAT;                      // :148
    const int fx = r.x + a / 5 - 2, fy = r.y + a % 5 - 2;                            // :151-152
    r.px = r.x; r.py = r.y; r.a = a;                                                 // this obs shows the old crop
    if (fx < W::G - 2 && fx > 1 && fy < W::G - 2 && fy > 1) { r.x = fx; r.y = fy; }  // :157-159
    else {                                                                           // :160-169 per-axis clamp
      if (fx >= W::G - 2) r.x = W::G - 3;
      if (fx <= 1) r.x = 2;
      if (fy >= W::G - 2) r.y = W::G - 3;
      if (fy <= 1) r.y = 2;
    }
    const int tc = t.cls[(r.L - 1) * W::G * W::G + fx * W::G + fy];
    if (fx == r.gx && fy == r.gy) { reward_code = RC_GOAL; o.cls = CLS_X; }          // :175-176
    else if (tc == CLS_W) { reward_code = RC_WALL; o.cls = CLS_W; }                  // :177-178
    else if (tc == CLS_B || tc == CLS_S) { reward_code = RC_MOVE; o.cls = CLS_B; }   // :179-180
    else o.cls = CLS_S;                                                              // 'X' not the goal: -0.0
    o.done = (reward_code == RC_GOAL) || (r.step > (uint32_t)W::STEP_LIMIT);        // :222
    p.reward[e] = __uint_as_float(reward_bits(reward_code));
    p.done[e] = o.done ? 1 : 0;
    if (o.done) o.eplen = r.step;
    reset_now = o.done && p.autoreset;
    o.render = true;
  } else if (p.mode == MODE_RESET) {
    reset_now = (p.mask == nullptr) || (p.mask[e] != 0);
    o.render = reset_now;
  } else {
    o.render = true;
  }
  if (reset_now) {
    uint32_t ep = p.episode[e];
    v2_respawn<W>(r, p, e, ep, t);
    p.episode[e] = ep;
    visit_op = 2;
  }
  out.info = (uint32_t)r.x | ((uint32_t)r.y << 5) | ((uint32_t)r.px << 10) | ((uint32_t)r.py << 15) | (visit_op << 20);
  uint32_t s, aux;
  v2_pack(r, s, aux);
  o.st = s;
  if (p.mode != MODE_RENDER) { p.state[e] = s; p.goal_count[e] = aux; }
  o.render = o.render && p.obs != nullptr && e >= p.win_lo && e < p.win_lo + p.win_n;
  out.mask[0] = v2_free_crop<W>(t, r.L, r.x, r.y);                                   // :185-186
  out.mask[1] = v2_goal_crop(r.x, r.y, r.gx, r.gy);
  out.mask[2] = r.a >= 0 ? (1u << r.a) : 0u;                                         // :136-137 / :87
  out.mask[3] = v2_free_crop<W>(t, r.L, r.px, r.py);                                 // retStatelast (:193)
  out.mask[4] = v2_goal_crop(r.px, r.py, r.gx, r.gy);
  return out;
}

// CTA-cooperative fused reset / step / render for the foveal variants (v2, v4).  Warp 0 grabs a 32-env
// tile from the global work counter, runs the transitions (one env per lane) and parks the five 25-bit
// planes of every env in shared memory; all threads then write the tile's contiguous bytes as float4
// stores (double-buffered: warp 0 is already on the next tile).  v4 adds one cooperative pass per tile
// over the envs' float visit layers: state[2] = (state[2] + visitMap) / 2 in float64, stored as float32
// (lmaze_env_v4.py:116-119,211-214), capturing the 5x5 crops at the ball and at the previous window.
template <class W, int THREADS>
__global__ void __launch_bounds__(THREADS) lmz_env_fov_kernel(const KParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t s_mask[2][32 * W::NBIT];
  __shared__ uint32_t s_info[2][32];
  __shared__ float s_vis[W::NVIS > 0 ? 2 * 32 * 50 : 1];   // [buf][env][0: at the ball, 1: at the previous window][25]
  __shared__ uint32_t s_flags[2];            // bit l: env l of the tile is to be rendered
  __shared__ long long s_tile[2];
  // v2: warp 0 produces (transitions) and then renders with everybody else.
  // v4: warps 0-3 are PRODUCERS -- warp 0 runs the next tile's transitions, then all four update that
  //     tile's float visit layers -- while the remaining warps render the current tile, so the visit
  //     pass (global read-modify-write latency) is hidden behind the obs stores.
  constexpr int PROD = (W::NVIS > 0) ? (THREADS >= 512 ? 128 : 64) : 0;   // threads that never render
  constexpr int CTHREADS = THREADS - PROD;
  static_assert(CTHREADS >= 32, "need at least one rendering warp");
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  stage_blob<W>(smem, &bar, p.blob);
  const FovTables<W> t(smem);
  const int64_t tiles = p.tile_end;
  WarpStats ws;

  auto produce = [&](int buf) {              // warp 0 only
    int64_t tl = 0;
    if (lane == 0) tl = p.tile_begin + grab_tile(p.work);
    tl = __shfl_sync(0xffffffffu, tl, 0);
    const int64_t e = tl * 32 + lane;
    const bool valid = tl < tiles && e < p.n;
    V2Lane v;
    v.o.st = 0; v.o.st_old = 0; v.o.render = false; v.o.done = false; v.o.cls = -1; v.o.eplen = 0; v.info = 0;
#pragma unroll
    for (int c = 0; c < W::NBIT; ++c) v.mask[c] = 0;
    if (valid) v = v2_lane<W>(p, e, t);
    if (p.mode == MODE_STEP) ws.add(valid, v.o);
#pragma unroll
    for (int c = 0; c < W::NBIT; ++c) s_mask[buf][lane * W::NBIT + c] = v.mask[c];
    s_info[buf][lane] = valid ? v.info : 0u;
    const unsigned fl = __ballot_sync(0xffffffffu, valid && v.o.render);
    if (lane == 0) { s_flags[buf] = fl; s_tile[buf] = tl; }
  };
  // v4 visit layers of one tile's 32 envs: 32 x 324 consecutive floats, coalesced read-modify-write by
  // the PROD producer threads: state[2] = (state[2] + visitMap) / 2 in float64 (lmaze_env_v4.py:211-214)
  auto visit_pass = [&](int buf) {
    const int64_t tile = s_tile[buf];
    if (tile >= tiles) return;
    const int64_t e0 = tile * 32;
    const uint32_t cells = (uint32_t)(((p.n - e0) < 32 ? (p.n - e0) : 32) * (W::G * W::G));
    float *vis = p.visit + e0 * (W::G * W::G);
    float *sv = s_vis + buf * (32 * 50);
    constexpr uint32_t NP = PROD > 0 ? PROD : 1, PER = (32 * W::G * W::G + NP - 1) / NP, UN = 8;
    for (uint32_t k0 = 0; k0 < PER; k0 += UN) {
      float vv[UN];
#pragma unroll
      for (uint32_t j = 0; j < UN; ++j) {                            // UN independent loads in flight per thread
        const uint32_t idx = tid + (k0 + j) * NP;
        vv[j] = (idx < cells) ? __ldcs(vis + idx) : 0.0f;
      }
#pragma unroll
      for (uint32_t j = 0; j < UN; ++j) {
        const uint32_t idx = tid + (k0 + j) * NP;
        if (idx >= cells) continue;
        const uint32_t env = idx / (W::G * W::G), cell = idx - env * (W::G * W::G);
        const int x = cell / W::G, y = cell - x * W::G;
        const uint32_t info = s_info[buf][env];
        const int bx = info & 31, by = (info >> 5) & 31, px = (info >> 10) & 31, py = (info >> 15) & 31;
        const uint32_t op = info >> 20;
        const int dx = x - bx + 2, dy = y - by + 2, qx = x - px + 2, qy = y - py + 2;
        const bool in_cur = dx >= 0 && dx < 5 && dy >= 0 && dy < 5;
        float v = vv[j];
        if (op == 1) v = (float)(((double)v + (in_cur ? 1.0 : 0.0)) * 0.5);
        else if (op == 2) v = in_cur ? 0.5f : 0.0f;                 // fresh zeros, then the same update (:106-113)
        if (op) __stcs(vis + idx, v);
        if (in_cur) sv[env * 50 + dx * 5 + dy] = v;
        if (qx >= 0 && qx < 5 && qy >= 0 && qy < 5) sv[env * 50 + 25 + qx * 5 + qy] = v;
      }
    }
  };
  auto producers_sync = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(PROD > 0 ? PROD : 32) : "memory"); };

  if (warp == 0) produce(0);
  if (W::NVIS > 0 && tid < PROD) { producers_sync(); visit_pass(0); }
  for (int buf = 0;; buf ^= 1) {
    __syncthreads();                          // tile(buf) is complete; buffers buf^1 are free again
    const int64_t tile = s_tile[buf];
    if (tile >= tiles) break;
    if (W::NVIS > 0 && tid < PROD) {
      if (warp == 0) produce(buf ^ 1);
      producers_sync();
      visit_pass(buf ^ 1);
      continue;
    }
    if (W::NVIS == 0 && warp == 0) produce(buf ^ 1);
    const uint32_t flags = s_flags[buf];
    if (flags == 0) continue;
    const int ctid = tid - PROD;
    const uint32_t *mk = s_mask[buf];
    const float *sv = s_vis + (W::NVIS > 0 ? buf * (32 * 50) : 0);
    const int64_t row0 = tile * 32 - p.win_lo;                     // obs row of the tile's first env
    float *dst = reinterpret_cast<float *>(p.obs) + row0 * (int64_t)W::OBS_FLOATS;
    auto value = [&](uint32_t env, uint32_t r) -> uint32_t {       // float bits of obs[env][r]
      const uint32_t code = t.lut[r], slot = code >> 5, cell = code & 31u;
      if (W::NVIS > 0 && slot >= (uint32_t)W::NBIT) return __float_as_uint(sv[env * 50 + (slot - W::NBIT) * 25 + cell]);
      return ((mk[env * W::NBIT + slot] >> cell) & 1u) ? 0x3f800000u : 0u;
    };
    if (flags == 0xffffffffu && (row0 & 3) == 0) {
      // fast path: the whole tile is rendered and 16-byte aligned -> TILE_F4 float4 stores
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
      // partial tile (batch tail, reset mask, render-window edge): guarded 32-bit stores
      for (uint32_t g = ctid; g < 32 * W::OBS_FLOATS; g += CTHREADS) {
        const uint32_t env = g / W::OBS_FLOATS, r = g - env * W::OBS_FLOATS;
        if ((flags >> env) & 1u) __stcs(reinterpret_cast<unsigned int *>(dst) + g, value(env, r));
      }
    }
  }
  if (warp == 0) {
    if (p.mode == MODE_STEP) ws.flush(p.stats, lane);
    if (lane == 0) finish_grabber(p.work, (unsigned long long)gridDim.x);
  }
}

// get/set_state for v2.  cols: x, y, goal_x, goal_y, step_count, layout, aux, episode
// aux = prev_x | prev_y << 5 | last_action << 10 | action_valid << 15
__global__ void lmz_state_v2_kernel(int64_t n, uint32_t *state, uint32_t *auxw, uint32_t *episode, int32_t *io, int set) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  int32_t *row = io + e * 8;
  if (set) {
    auto clampi = [](int v) { return v < 2 ? 2 : (v > V2::G - 3 ? V2::G - 3 : v); };   // ball stays in [2, 15]
    V2Regs r;
    r.x = clampi(row[0]); r.y = clampi(row[1]); r.gx = row[2] & 31; r.gy = row[3] & 31;
    r.step = (uint32_t)row[4] < V2::STEP_SAT ? (uint32_t)row[4] : V2::STEP_SAT;
    r.L = row[5] < 1 ? 1 : (row[5] > 5 ? 5 : row[5]);
    r.px = clampi(row[6] & 31); r.py = clampi((row[6] >> 5) & 31);
    r.a = ((row[6] >> 15) & 1) ? ((row[6] >> 10) & 31) : -1;
    if (r.a > 24) r.a = 24;
    uint32_t s, aux;
    v2_pack(r, s, aux);
    state[e] = s; auxw[e] = aux; episode[e] = (uint32_t)row[7];
  } else {
    const V2Regs r = v2_unpack(state[e], auxw[e]);
    row[0] = r.x; row[1] = r.y; row[2] = r.gx; row[3] = r.gy; row[4] = (int32_t)r.step; row[5] = r.L;
    row[6] = (int32_t)auxw[e]; row[7] = (int32_t)episode[e];
  }
}

}  // namespace lmz
