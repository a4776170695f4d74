// lmz_fov_rollout.cuh -- T fused steps of the FOVEAL variants with no per-step observation:
// lmaze-v2 / lmaze-v4 (lmaze_env_v2.py:127-225, lmaze_env_v4.py:155-275) and the planner / actor env lmaze-v5 / v6
// (lmaze_env_v5.py:158-292), the counterpart of lmz_rollout_kernel (v0 / v3) in lmz_kernels.cuh.
//
// One thread per env, the packed state words in registers for all T steps, the per-maze tables (cell classes,
// spawn candidate lists) in shared memory, per-step reward f32 (or 1-byte reward code) and done u8 [T][N] streaming
// stores, coalesced across envs.  Actions come from a [T][N] device buffer or from Philox.  The same transition
// functions as the fused step kernel run here (v2_step_core, v5_planner_core, v5_step_core, v2_respawn): there is
// one statement of the env logic.  v4 / v5 keep their float visit layer in HBM in the SCALED form (lmz_v2.cuh): an
// averaging step is 25 read-modify-writes on the window around the ball, by the env's own thread.
//
// v5 / v6: one rollout step = plannerStep for the envs that are waiting for their planner (localDone set, or no
// plannerStep since the last reset -- the device-side mask of lmz_planner_step_auto) followed by step(); outputs
// are globalReward, originalReward, globalDone, localDone per step.
#pragma once
#include "lmz_fov.cuh"

namespace lmz {

constexpr uint32_t TAG_ACTION25 = 0x42u, TAG_HIER = 0x48u;

// the visit layer of ONE env in GLOBAL memory handled by ONE thread (short rollouts)
template <class W>
__device__ __forceinline__ void visit_thread(float *vis, uint32_t vinfo, int bx, int by) {
  const uint32_t op = vinfo & 7u;
  if (op == VOP_READ) return;
  const int tpre = (vinfo >> 3) & 127;
  if (op == VOP_AVG) {                                     // s' = RN(s + 2^T) on the 5x5 window
    const float add = __int_as_float((127 + tpre) << 23);
    float *w = vis + (bx - 2) * W::G + (by - 2);
    float v[25];
#pragma unroll
    for (int k = 0; k < 25; ++k) v[k] = __ldcg(w + (k / 5) * W::G + k % 5);
#pragma unroll
    for (int k = 0; k < 25; ++k) __stcg(w + (k / 5) * W::G + k % 5, __fadd_rn(v[k], add));
    return;
  }
  const float down = visit_scale_down(tpre);               // reset, or the literal full pass of direct mode
  float4 *row = reinterpret_cast<float4 *>(vis);
  for (int c4 = 0; c4 < W::G * W::G / 4; ++c4) {
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
    if (op == VOP_FULL) q = __ldcg(row + c4);
    float *f = reinterpret_cast<float *>(&q);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int cell = 4 * c4 + k, x = cell / W::G, y = cell - x * W::G;
      const bool in_cur = (unsigned)(x - bx + 2) < 5u && (unsigned)(y - by + 2) < 5u;
      f[k] = (op == VOP_FULL) ? visit_average(__fmul_rn(f[k], down), in_cur) : W::visit_reset_stored(in_cur);
    }
    __stcg(row + c4, q);
  }
}

// The shared-memory form used by long rollouts.  A reset must not cost 324 stores -- in a warp of 32 envs SOME lane
// resets in every other step, and the whole warp would walk the loop -- so rows are zeroed LAZILY: `rowmask` (bit r =
// row r of the copy holds valid data) is cleared by a reset, a row is zero-filled the first time a window touches it,
// and whatever is still unmaterialised at the end of the launch is zero-filled before the layer is copied out.
template <class W>
__device__ __forceinline__ void visit_rows_valid(float *vis, uint32_t &rowmask, uint32_t want) {
  uint32_t miss = want & ~rowmask;
  while (miss) {
    const int r = __ffs(miss) - 1;
    miss &= miss - 1;
#pragma unroll
    for (int y = 0; y < W::G; ++y) vis[r * W::G + y] = 0.0f;
  }
  rowmask |= want;
}
template <class W>
__device__ __forceinline__ void visit_thread_smem(float *vis, uint32_t vinfo, int bx, int by, uint32_t &rowmask) {
  const uint32_t op = vinfo & 7u;
  if (op == VOP_READ) return;
  const int tpre = (vinfo >> 3) & 127;
  float *w = vis + (bx - 2) * W::G + (by - 2);
  if (op == VOP_RESET) {                                   // zeros everywhere (lazily), the reset value on the window
    rowmask = 0;
    visit_rows_valid<W>(vis, rowmask, 31u << (bx - 2));
    const float v = W::visit_reset_stored(true);
    if (v != 0.0f) {
#pragma unroll
      for (int k = 0; k < 25; ++k) w[(k / 5) * W::G + k % 5] = v;
    }
    return;
  }
  if (op == VOP_AVG) {                                     // s' = RN(s + 2^T) on the 5x5 window
    visit_rows_valid<W>(vis, rowmask, 31u << (bx - 2));
    const float add = __int_as_float((127 + tpre) << 23);
#pragma unroll
    for (int k = 0; k < 25; ++k) w[(k / 5) * W::G + k % 5] = __fadd_rn(w[(k / 5) * W::G + k % 5], add);
    return;
  }
  visit_rows_valid<W>(vis, rowmask, (1u << W::G) - 1u);    // direct mode: the literal full pass
  const float down = visit_scale_down(tpre);
  for (int cell = 0; cell < W::G * W::G; ++cell) {
    const int x = cell / W::G, y = cell - x * W::G;
    const bool in_cur = (unsigned)(x - bx + 2) < 5u && (unsigned)(y - by + 2) < 5u;
    vis[cell] = visit_average(__fmul_rn(vis[cell], down), in_cur);
  }
}

// Long rollouts keep every env's layer in SHARED memory for the whole launch: a window update per step in global memory
// is 50 uncoalesced accesses per thread (32 different sectors per warp instruction; v4, 2^21 envs x T=64: 69.7 ms), in
// shared memory it is 50 conflict-free LDS / STS.  The CTA's layers (consecutive envs: one contiguous block) are copied
// in and out with coalesced accesses, 1,296 B read + 1,296 B written per env per LAUNCH.
constexpr int VIS_STRIDE = 325;                            // floats between two threads' layers (odd: no bank conflicts)
constexpr int VIS_SMEM_MIN_T = 8;                          // shorter rollouts update the layer in global memory

struct RolloutCounters {
  uint32_t steps = 0, ep = 0, goal = 0, wall = 0, move = 0, stale = 0;
  unsigned long long len = 0;
};

__device__ __forceinline__ void rollout_store_reward(const KParams &p, int64_t off, int rcode) {
  if (p.reward_code) __stcs(p.reward_code + off, (uint8_t)rcode);
  else __stcs(p.reward + off, __uint_as_float(reward_bits(rcode)));
}

// ---- lmaze-v2 / lmaze-v4 ------------------------------------------------------------------------------------
// Philox action of env g at global rollout step t: word t & 3 of philox(ctr = (g_lo, g_hi, b_lo, (b_hi << 8) | 0x42)),
// b = t >> 2; action = (word * 25) >> 32.
template <class W, bool VIS_SMEM>
__device__ __forceinline__ void rollout_env_v2(const KParams &p, int64_t e, const FovTables<W> &t, RolloutCounters &c,
                                               float *vis) {
  V2Regs r = v2_unpack(p.state[e], p.goal_count[e]);
  uint32_t ep = p.episode[e];
  const uint64_t gid = p.env_id0 + (uint64_t)e;
  uint32_t w[4] = {0, 0, 0, 0};
  uint32_t rowmask = (1u << W::G) - 1u;                    // VIS_SMEM: every row of the copy is valid at the start
  uint64_t tg = p.t0;
  int64_t off = e;
  for (int s = 0; s < p.T; ++s, ++tg, off += p.n) {
    long long a;
    if (p.actions) a = load_action(p.actions, p.action_dtype, off);
    else {
      if ((tg & 3) == 0 || s == 0) {
        const uint64_t blk = tg >> 2;
        philox4x32_10((uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)blk, ((uint32_t)(blk >> 32) << 8) | TAG_ACTION25,
                      (uint32_t)p.seed, (uint32_t)(p.seed >> 32), w);
      }
      const uint32_t k = (uint32_t)(tg & 3);
      a = __umulhi(k == 0 ? w[0] : k == 1 ? w[1] : k == 2 ? w[2] : w[3], 25u);
    }
    int rcode, cls = -1;
    const bool done = v2_step_core<W>(r, a, t, p.errors, rcode, cls);
    rollout_store_reward(p, off, rcode);
    __stcs(p.done + off, (uint8_t)(done ? 1 : 0));
    c.wall += (cls == CLS_W); c.move += (cls == CLS_B || cls == CLS_X); c.stale += (cls == CLS_S);
    int want = 1;                                          // v4: every step averages the layer (lmaze_env_v4.py:211-214)
    if (done) {
      c.ep += 1; c.len += r.step; c.goal += (cls == CLS_X);
      if (p.autoreset) { v2_respawn<W>(r, p, e, ep, t); want = 2; }
    }
    if (W::NVIS > 0) {
      const uint32_t vinfo = visit_plan<W>(want, r.vt);
      if (VIS_SMEM) visit_thread_smem<W>(vis, vinfo, r.x, r.y, rowmask); else visit_thread<W>(vis, vinfo, r.x, r.y);
    }
  }
  if (W::NVIS > 0 && VIS_SMEM) visit_rows_valid<W>(vis, rowmask, (1u << W::G) - 1u);   // materialise the lazy zeros
  c.steps += (uint32_t)p.T;
  uint32_t w0, w1;
  v2_pack(r, w0, w1);
  p.state[e] = w0; p.goal_count[e] = w1; p.episode[e] = ep;
}

// ---- lmaze-v5 / lmaze-v6 ------------------------------------------------------------------------------------
// Philox goal / action of env g at global rollout step t: philox(ctr = (g_lo, g_hi, t_lo, (t_hi << 8) | 0x48));
// goal = (word0 * 25) >> 32, action = word1 >> 30.
template <class W, bool VIS_SMEM>
__device__ __forceinline__ void rollout_env_v5(const KParams &p, int64_t e, const FovTables<W> &t, const unsigned char *sb,
                                               RolloutCounters &c, float *vis) {
  V5Regs r = v5_unpack(p.state[e], p.goal_count[e], p.aux2[e]);
  uint32_t ep = p.episode[e];
  const uint64_t gid = p.env_id0 + (uint64_t)e;
  uint32_t rowmask = (1u << W::G) - 1u;                    // VIS_SMEM: every row of the copy is valid at the start
  uint64_t tg = p.t0;
  int64_t off = e;
  for (int s = 0; s < p.T; ++s, ++tg, off += p.n) {
    long long a, g;
    if (p.actions) {
      a = load_action(p.actions, p.action_dtype, off);
      g = load_action(p.goals, p.action_dtype, off);
    } else {
      uint32_t w[4];
      philox4x32_10((uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)tg, ((uint32_t)(tg >> 32) << 8) | TAG_HIER,
                    (uint32_t)p.seed, (uint32_t)(p.seed >> 32), w);
      g = __umulhi(w[0], 25u);
      a = w[1] >> 30;
    }
    if (r.ld != 0 || r.fstep == 0) v5_planner_core<W>(r, g, p.errors);      // plannerStep for the envs waiting for it
    int lcode, gcode, cls = -1;
    const bool done = v5_step_core<W>(r, a, t, lcode, gcode, cls);
    int want = r.ld ? 1 : 0;                                                 // lmaze_env_v5.py:308-312
    if (r.fstep == 0 || r.ld) { r.lx = r.x; r.ly = r.y; }                    // :327-328, :351-352
    __stcs(p.reward + off, __uint_as_float(reward_bits(gcode)));
    __stcs(p.reward2 + off, __uint_as_float(reward_bits(lcode)));
    __stcs(p.done + off, (uint8_t)r.gd);
    __stcs(p.done2 + off, (uint8_t)r.ld);
    c.wall += (cls == CLS_W); c.move += (cls == CLS_B || cls == CLS_X);
    if (done) { c.ep += 1; c.len += r.fstep; c.goal += (cls == CLS_X); }
    if (r.gd && p.autoreset) { v5_respawn<W>(r, p, e, ep, t, sb); want = 2; }
    const uint32_t vinfo = visit_plan<W>(want, r.vt);
    if (VIS_SMEM) visit_thread_smem<W>(vis, vinfo, r.x, r.y, rowmask); else visit_thread<W>(vis, vinfo, r.x, r.y);
  }
  if (VIS_SMEM) visit_rows_valid<W>(vis, rowmask, (1u << W::G) - 1u);                  // materialise the lazy zeros
  c.steps += (uint32_t)p.T;
  uint32_t w0, w1, w2;
  v5_pack(r, w0, w1, w2);
  p.state[e] = w0; p.goal_count[e] = w1; p.aux2[e] = w2; p.episode[e] = ep;
  if (p.fgoal_out) p.fgoal_out[e] = (uint8_t)r.fga;
}

template <class W, int THREADS, bool VIS_SMEM>
__global__ void __launch_bounds__(THREADS) lmz_fov_rollout_kernel(const KParams p) {
  extern __shared__ __align__(128) unsigned char smem[];         // the per-maze tables (blob from ROWBITS_OFF on) [+ the layers]
  __shared__ unsigned long long blk_stats[NUM_STATS];
  constexpr uint32_t STAGE_OFF = W::ROWBITS_OFF;
  constexpr uint32_t TAB_BYTES = (W::BLOB_BYTES - STAGE_OFF + 15u) & ~15u;
  constexpr int GG = W::G * W::G;
  for (uint32_t i = threadIdx.x; i < (W::BLOB_BYTES - STAGE_OFF) / 4; i += THREADS)
    reinterpret_cast<uint32_t *>(smem)[i] = reinterpret_cast<const uint32_t *>(p.blob + STAGE_OFF)[i];
  if (threadIdx.x < NUM_STATS) blk_stats[threadIdx.x] = 0;
  __syncthreads();
  const unsigned char *sb = smem - STAGE_OFF;
  const FovTables<W> t(sb);
  float *layers = reinterpret_cast<float *>(smem + TAB_BYTES);   // VIS_SMEM: [THREADS][VIS_STRIDE]
  RolloutCounters c;
  for (int64_t base = (int64_t)blockIdx.x * THREADS; base < p.n; base += (int64_t)gridDim.x * THREADS) {
    const int64_t e = base + threadIdx.x;
    const int cnt = (int)((p.n - base) < THREADS ? (p.n - base) : THREADS);        // envs of this CTA pass
    float *vis = nullptr;
    if (W::NVIS > 0) {
      if (VIS_SMEM) {                                            // the pass's layers are one contiguous block: coalesced copy in
        const float *src = p.visit + base * GG;
        for (int i = threadIdx.x; i < cnt * GG; i += THREADS) layers[(i / GG) * VIS_STRIDE + i % GG] = __ldcs(src + i);
        __syncthreads();
        vis = layers + threadIdx.x * VIS_STRIDE;
      } else {
        vis = p.visit + e * GG;
      }
    }
    if (e < p.n) {
      if constexpr (W::HAS_LOC) rollout_env_v5<W, VIS_SMEM>(p, e, t, sb, c, vis);
      else rollout_env_v2<W, VIS_SMEM>(p, e, t, c, vis);
    }
    if (W::NVIS > 0 && VIS_SMEM) {                               // ... and out
      __syncthreads();
      float *dst = p.visit + base * GG;
      for (int i = threadIdx.x; i < cnt * GG; i += THREADS) __stcs(dst + i, layers[(i / GG) * VIS_STRIDE + i % GG]);
      __syncthreads();
    }
  }
  unsigned long long v[NUM_STATS] = {c.steps, c.ep, c.goal, (unsigned long long)(c.ep - c.goal), c.wall, c.move, c.stale, c.len};
#pragma unroll
  for (int k = 0; k < NUM_STATS; ++k) {
    unsigned long long x = v[k];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) x += __shfl_down_sync(0xffffffffu, x, off);
    if ((threadIdx.x & 31) == 0 && x) atomicAdd(&blk_stats[k], x);
  }
  __syncthreads();
  if (threadIdx.x < NUM_STATS && blk_stats[threadIdx.x]) atomicAdd(&p.stats[threadIdx.x], blk_stats[threadIdx.x]);
}

}  // namespace lmz
