// lmz_fov.cuh -- the fused reset / step / render kernel of the FOVEAL variants: lmaze-v2, lmaze-v4
// (lmz_v2.cuh) and the planner / actor env lmaze-v5 / v6 (lmz_v5.cuh, which adds plannerStep and a second,
// "local" observation tensor).
//
// An observation here is C planes of 5x5 values, each upsampled x7 to 35x35 floats
// (lmaze_env_v2.py:197-203, lmaze_env_v4.py:241-247, lmaze_env_v5.py:337-343,372-378): 6,125 / 8,575
// floats per env, NOT a multiple of 4, so an env row is not 16-byte aligned -- but four consecutive rows
// are.  The render therefore works on GROUPS of 4 envs: a group's image is OBS_FLOATS float4s, and because
// every cell value repeats 7 times along a row, the four floats of one float4 take at most TWO different
// cell values (a prefix of k floats from cell A, the rest from cell B).  One host-built table entry per
// float4 -- A:11 | B:11 | k:3, A and B indexing the group's [4][NSLOT][25] value planes -- turns the whole
// render into: 1 table load, 2 value loads, 3 selects, 1 streaming 128-bit store.  (The first version looked
// every FLOAT up through a byte table and was issue-bound at 91 % issue-slot utilisation, 5.8 TB/s on v4.)
//
// CTA organisation: ONE small persistent CTA per SM; tiles of 32 envs are handed out by the global work counter.
// Warp 0 is a dedicated producer: it runs the tile's 32 transitions (one env per lane, state in registers) and
// expands each env's 25-bit planes into float planes in shared memory, one tile ahead of the rendering warps and
// one tile ahead of its OWN loads (next tile's index, state words, actions).  For the variants with a float visit
// layer (v4, v5) warps 0-3 are producers: the tile's 32 x 324 visit floats arrive by ONE bulk async (TMA) load,
// the producers update them out of shared memory (coalesced write-back) and drop the two 5x5 visit crops into the
// same planes.  Everything is double buffered; one __syncthreads per tile.  FEWER rendering warps reach a HIGHER
// write bandwidth (tools/fov_sweep2.py): v2 runs 1 + 3 warps per SM (7.47 TB/s, the pure-write ceiling; 7.0 TB/s
// with 32 warps), v4 4 + 3, v5 4 + 4.
#pragma once
#include "lmz_v2.cuh"
#include "lmz_v5.cuh"

#ifndef LMZ_VISIT_PROD
#define LMZ_VISIT_PROD 128
#endif
#ifndef LMZ_V2_PROD
#define LMZ_V2_PROD 32    // v2: one dedicated producer warp (0: warp 0 produces, then renders with the others)
#endif

namespace lmz {

// state[2] = (state[2] + visitMap) / 2, which the reference evaluates in float64 and stores as float32
// (lmaze_env_v4.py:116-119,211-214; lmaze_env_v5.py:308-312).  For 0 <= v <= 1 the float32 expression below is
// bit-identical to float32((float64(v) + m) * 0.5): v/2 is a single correctly rounded operation either way, and
// v + 1 is either exact in float64 (v >= 2^-29) -- then both paths round the same exact value at the same bit --
// or so close to 1 that both give exactly 1.0 before the halving.  (Checked bit for bit by the v4 / v5 parity
// tests over the whole layer.)  It keeps FP64 and the f32<->f64 conversions off the producers' path.
__device__ __forceinline__ float visit_average(float v, bool in_window) {
  return in_window ? __fmul_rn(__fadd_rn(v, 1.0f), 0.5f) : __fmul_rn(v, 0.5f);
}

template <int ID, int C>
__device__ __forceinline__ FovLane<5> fov_lane(const Fov<ID, C> *, const KParams &p, int64_t e,
                                               const FovTables<Fov<ID, C>> &t, const unsigned char *, const FovPre &pre) {
  return v2_lane<Fov<ID, C>>(p, e, t, pre);
}
__device__ __forceinline__ V5Lane fov_lane(const V5 *, const KParams &p, int64_t e, const FovTables<V5> &t,
                                           const unsigned char *smem, const FovPre &pre) {
  return v5_lane<V5>(p, e, t, smem, pre);
}

template <class W, int THREADS>
__global__ void __launch_bounds__(THREADS) lmz_env_fov_kernel(const KParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ __align__(8) uint64_t vbar[2];
  __shared__ uint32_t s_info[2][32];
  __shared__ uint32_t s_flags[2][2];         // [buf][0 obs, 1 local obs]: bit l = env l of the tile is written
  __shared__ long long s_tile[2];
  constexpr int PROD = (W::NVIS > 0) ? LMZ_VISIT_PROD : LMZ_V2_PROD;   // threads that never render
  constexpr int CTHREADS = THREADS - PROD;
  static_assert(CTHREADS >= 32, "need at least one rendering warp");
  static_assert((uint32_t)CTHREADS < W::OBS_FLOATS, "index stepping assumes fewer threads than entries");
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (W::NVIS > 0 && tid == 0) { mbar_init(&vbar[0], 1); mbar_init(&vbar[1], 1); }   // fenced + synced inside stage_blob
  stage_blob<W>(smem, &bar, p.blob);
  const FovTables<W> t(smem);
  float *vals = reinterpret_cast<float *>(smem + W::BLOB_BYTES);          // [2][32][NSLOT][25]
  float *vbuf = vals + 2 * 32 * W::VALS;                                  // [2][32][324]: TMA landing zone of the visit layers
  uint32_t vphase = 0;                                                    // bit b: parity the next wait on vbar[b] expects
  const int64_t tiles = p.tile_end;
  const bool need_visit = W::NVIS > 0 && p.mode != MODE_PLANNER;
  WarpStats ws;

  // Warp 0 runs ONE TILE AHEAD of its own loads: when it turns to a tile, the tile's index was grabbed and its
  // state words / actions (and, for the visit variants, the bulk load of its 32 visit layers) were issued a
  // whole tile time earlier, so no DRAM or atomic round trip -- several microseconds each under a saturated
  // write stream -- sits between two tiles of the producer.
  int64_t tl_next = 0;                       // warp 0: the tile produce() turns to next
  FovPre pre_next;                           // ... and its preloaded words (this lane's env)
  pre_next.w0 = pre_next.w1 = pre_next.w2 = 0u; pre_next.act = 0;
  auto prefetch = [&](int64_t tl, int vb) {  // warp 0: start everything tile `tl` will need
    if (W::NVIS > 0 && need_visit && lane == 0 && tl < tiles) {
      // the tile's 32 visit layers are 41,472 contiguous bytes: ONE bulk async (TMA) load
      const int64_t e0 = tl * 32;
      const uint32_t bytes = (uint32_t)(((p.n - e0) < 32 ? (p.n - e0) : 32) * (W::G * W::G * 4));
      mbar_expect_tx(&vbar[vb], bytes);
      bulk_g2s(vbuf + vb * (32 * W::G * W::G), p.visit + e0 * (W::G * W::G), bytes, &vbar[vb]);
    }
    const int64_t e = tl * 32 + lane;
    if (tl < tiles && e < p.n) pre_next = fov_preload<W>(p, e);
    tl_next = tl;
  };
  auto grab = [&]() {
    int64_t tl = 0;
    if (lane == 0) tl = p.tile_begin + grab_tile(p.work);
    return __shfl_sync(0xffffffffu, tl, 0);
  };
  auto produce = [&](int buf) {              // warp 0 only
    const int64_t tl = tl_next;
    const FovPre pre = pre_next;
    const int64_t tl_after = grab();         // in flight while this tile's transitions run
    const int64_t e = tl * 32 + lane;
    const bool valid = tl < tiles && e < p.n;
    FovLane<W::NBIT> v;
    v.o.st = 0; v.o.st_old = 0; v.o.render = false; v.o.done = false; v.o.cls = -1; v.o.eplen = 0; v.info = 0;
    v.rfov = false; v.rloc = false;
    if (valid) v = fov_lane(static_cast<const W *>(nullptr), p, e, t, smem, pre);
    if (p.mode == MODE_STEP) ws.add(valid, v.o);
    if (valid && (v.rfov || v.rloc)) {       // 25-bit planes -> float planes (lane stride VALS is odd: no bank conflicts)
      float *mv = vals + (buf * 32 + lane) * W::VALS;
#pragma unroll
      for (int b = 0; b < W::NBIT; ++b) {
        const uint32_t m = v.mask[b];
        float *pl = mv + W::bit_slot(b) * 25;
#pragma unroll
        for (int c = 0; c < 25; ++c) pl[c] = ((m >> c) & 1u) ? 1.0f : 0.0f;
      }
    }
    s_info[buf][lane] = valid ? v.info : 0u;
    const unsigned ff = __ballot_sync(0xffffffffu, valid && v.rfov);
    const unsigned fl = __ballot_sync(0xffffffffu, valid && v.rloc);
    if (lane == 0) { s_flags[buf][0] = ff; s_flags[buf][1] = fl; s_tile[buf] = tl; }
    prefetch(tl_after, buf ^ 1);             // vbuf[buf ^ 1] is free: its tile's visit pass ended before the last barrier
  };
  // visit layers of one tile's 32 envs: 32 x 324 consecutive floats, coalesced pass by the PROD producer
  // threads: state[2] = (state[2] + visitMap) / 2 in float64, stored as float32
  // (lmaze_env_v4.py:116-119,211-214; lmaze_env_v5.py:308-312)
  auto visit_pass = [&](int buf) {
    const int64_t tile = s_tile[buf];
    if (tile >= tiles || !need_visit) return;
    const int64_t e0 = tile * 32;
    const uint32_t cells = (uint32_t)(((p.n - e0) < 32 ? (p.n - e0) : 32) * (W::G * W::G));
    float *vis = p.visit + e0 * (W::G * W::G);
    float *tv = vals + buf * 32 * W::VALS;
    const float *vb = vbuf + buf * (32 * W::G * W::G);
    mbar_wait(&vbar[buf], (vphase >> buf) & 1u);
    vphase ^= 1u << buf;
    constexpr uint32_t NP = PROD > 0 ? PROD : 1;
#pragma unroll 4
    for (uint32_t idx = tid; idx < cells; idx += NP) {
      const uint32_t env = idx / (W::G * W::G), cell = idx - env * (W::G * W::G);
      const int x = cell / W::G, y = cell - x * W::G;
      const uint32_t info = s_info[buf][env];
      const int bx = info & 31, by = (info >> 5) & 31, px = (info >> 10) & 31, py = (info >> 15) & 31;
      const uint32_t op = (info >> 20) & 3u;
      const int dx = x - bx + 2, dy = y - by + 2, qx = x - px + 2, qy = y - py + 2;
      const bool in_cur = dx >= 0 && dx < 5 && dy >= 0 && dy < 5;
      float v = vb[idx];
      if (op == 1) v = visit_average(v, in_cur);
      else if (op == 2) v = W::visit_reset(in_cur);
      if (op) __stcs(vis + idx, v);
      if (in_cur) tv[env * W::VALS + W::VIS_SLOT0 * 25 + dx * 5 + dy] = v;
      if (qx >= 0 && qx < 5 && qy >= 0 && qy < 5) tv[env * W::VALS + W::VIS_SLOT1 * 25 + qx * 5 + qy] = v;
    }
  };
  auto producers_sync = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(PROD > 0 ? PROD : 32) : "memory"); };
  auto pick = [](uint32_t en, const float *gv) -> uint4 { return f4_pick(en, gv); };

  if (warp == 0) { prefetch(grab(), 0); produce(0); }
  if (W::NVIS > 0 && tid < PROD) { producers_sync(); visit_pass(0); }
  for (int buf = 0;; buf ^= 1) {
    __syncthreads();                          // tile(buf) is complete; buffers buf^1 are free again
    const int64_t tile = s_tile[buf];
    if (tile >= tiles) break;
    if (PROD > 0 && tid < PROD) {
      if (warp == 0) produce(buf ^ 1);
      if (W::NVIS > 0) { producers_sync(); visit_pass(buf ^ 1); }
      continue;
    }
    if (PROD == 0 && warp == 0) produce(buf ^ 1);
    const int ctid = tid - PROD;
    const float *tv = vals + buf * 32 * W::VALS;
    const int64_t row0 = tile * 32 - p.win_lo;                     // obs row of the tile's first env
    uint32_t flags = s_flags[buf][0];
    if (flags) {
      unsigned char *dst = reinterpret_cast<unsigned char *>(p.obs) + row0 * (int64_t)W::OBS_BYTES;
      constexpr uint32_t TOTAL = 8 * W::OBS_FLOATS;                // float4s of the tile = 8 groups x OBS_FLOATS entries
      uint32_t ent = ctid, grp = 0;
      if (flags == 0xffffffffu && (row0 & 3) == 0) {
        // fast path: the whole tile is rendered and 16-byte aligned
#pragma unroll 4
        for (uint32_t q = ctid; q < TOTAL; q += CTHREADS) {
          st_stream_v4(dst + ((size_t)q << 4), pick(t.lut[ent], tv + grp * 4 * W::VALS));
          ent += CTHREADS;
          if (ent >= W::OBS_FLOATS) { ent -= W::OBS_FLOATS; ++grp; }
        }
      } else {
        // partial tile (batch tail, reset mask, render-window edge): guarded 32-bit stores
        for (uint32_t q = ctid; q < TOTAL; q += CTHREADS) {
          const uint32_t en = t.lut[ent], a = en & 2047u, b = (en >> 11) & 2047u, k = en >> 22;
          const float *gv = tv + grp * 4 * W::VALS;
#pragma unroll
          for (uint32_t j = 0; j < 4; ++j) {
            const uint32_t code = j < k ? a : b, env = grp * 4 + code / W::VALS;
            if ((flags >> env) & 1u)
              __stcs(reinterpret_cast<unsigned int *>(dst) + ((size_t)q * 4 + j), __float_as_uint(gv[code]));
          }
          ent += CTHREADS;
          if (ent >= W::OBS_FLOATS) { ent -= W::OBS_FLOATS; ++grp; }
        }
      }
    }
    if (W::HAS_LOC) {
      // local observation (lmaze_env_v5.py:356-380): every env row is 16-byte aligned (19,600 = 16 x 1,225)
      flags = s_flags[buf][1];
      if (flags) {
        constexpr uint32_t PER = W::LOC_FLOATS / 4 > 0 ? W::LOC_FLOATS / 4 : 1;
        const uint32_t *loclut = reinterpret_cast<const uint32_t *>(smem + W::LOCLUT_OFF);
        unsigned char *dst = reinterpret_cast<unsigned char *>(p.obs2) + row0 * (int64_t)(W::LOC_FLOATS * 4);
        if (flags == 0xffffffffu) {
          uint32_t ent = ctid % PER, env = ctid / PER;
#pragma unroll 4
          for (uint32_t q = ctid; q < 32 * PER; q += CTHREADS) {
            // (an IndexError row is all zero: the producer zeroed the four local planes of that env)
            st_stream_v4(dst + ((size_t)q << 4), pick(loclut[ent], tv + env * W::VALS));
            ent += CTHREADS;
            while (ent >= PER) { ent -= PER; ++env; }
          }
        } else {
          while (flags) {                                          // a few envs of the tile (masked plannerStep)
            const uint32_t env = __ffs(flags) - 1;
            flags &= flags - 1;
            for (uint32_t q = ctid; q < PER; q += CTHREADS)
              st_stream_v4(dst + ((size_t)(env * PER + q) << 4), pick(loclut[q], tv + env * W::VALS));
          }
        }
      }
    }
  }
  if (warp == 0) {
    if (p.mode == MODE_STEP) ws.flush(p.stats, lane);
    if (lane == 0) finish_grabber(p.work, (unsigned long long)gridDim.x);
  }
}

// ---- compact observations / transition-only launches of the foveal variants -----------------------------------
// obs_mode = LMZ_OBS_COMPACT: the observation is written as f32 [C][5][5] -- the reference's `retState` BEFORE its
// x7 upsample (lmaze_env_v2.py:185-203, lmaze_env_v4.py:236-247, lmaze_env_v5.py:333-343,372-378) -- 500 / 700 /
// 700 + 400 bytes per env instead of 24,500 / 34,300 / 53,900; the reference image is exactly
// repeat_interleave(compact, 7) on both axes.  The same kernel serves launches with no observation bound.  The work
// per tile is small, so the organisation is warp-granular like lmz_env_compact_kernel: every warp grabs its own
// tiles (the next tile's index and state words requested one tile ahead), runs the 32 transitions, makes the
// coalesced pass over the tile's visit layers itself (v4 / v5) and stores the tile's planes as one flat run of
// coalesced 32-bit words.
template <class W, int THREADS>
__global__ void __launch_bounds__(THREADS) lmz_fov_small_kernel(const KParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  constexpr int WARPS = THREADS / 32;
  constexpr uint32_t STAGE_OFF = W::ROWBITS_OFF;                 // the render tables are not needed here
  __shared__ uint32_t s_mask[WARPS][32 * W::NSLOT];          // 25-bit planes indexed by value-plane slot
  __shared__ uint32_t s_info[WARPS][32];
  __shared__ float s_crop[W::NVIS > 0 ? WARPS : 1][W::NVIS > 0 ? 32 * 50 : 1];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(&bar, W::BLOB_BYTES - STAGE_OFF);
    bulk_g2s(smem, p.blob + STAGE_OFF, W::BLOB_BYTES - STAGE_OFF, &bar);
  }
  mbar_wait(&bar, 0);
  const unsigned char *sb = smem - STAGE_OFF;                    // sb + X_OFF addresses staged table X
  const FovTables<W> t(sb);
  const int64_t tiles = p.tile_end;
  const bool need_visit = W::NVIS > 0 && p.mode != MODE_PLANNER;
  uint32_t *mk = s_mask[warp];
  WarpStats ws;
  auto grab = [&]() {
    int64_t tl = 0;
    if (lane == 0) tl = p.tile_begin + grab_tile(p.work);
    return __shfl_sync(0xffffffffu, tl, 0);
  };
  auto preload = [&](int64_t tl, FovPre &q) {
    const int64_t e = tl * 32 + lane;
    q.w0 = q.w1 = q.w2 = 0u; q.act = 0;
    if (tl < tiles && e < p.n) q = fov_preload<W>(p, e);
  };
  int64_t tile = grab(), ntile = grab();
  FovPre pre;
  preload(tile, pre);
  while (tile < tiles) {
    const int64_t nntile = grab();                               // needed two tiles from now
    const int64_t e = tile * 32 + lane;
    const bool valid = e < p.n;
    FovLane<W::NBIT> v;
    v.o.st = 0; v.o.st_old = 0; v.o.render = false; v.o.done = false; v.o.cls = -1; v.o.eplen = 0; v.info = 0;
    v.rfov = false; v.rloc = false;
#pragma unroll
    for (int b = 0; b < W::NBIT; ++b) v.mask[b] = 0;
    if (valid) v = fov_lane(static_cast<const W *>(nullptr), p, e, t, sb, pre);
    if (p.mode == MODE_STEP) ws.add(valid, v.o);
    preload(ntile, pre);                                         // next tile's words fly while this tile is written
    const unsigned ff = __ballot_sync(0xffffffffu, valid && v.rfov);
    const unsigned fl = __ballot_sync(0xffffffffu, valid && v.rloc);
    __syncwarp();
#pragma unroll
    for (int b = 0; b < W::NBIT; ++b) mk[lane * W::NSLOT + W::bit_slot(b)] = v.mask[b];
    s_info[warp][lane] = valid ? v.info : 0u;
    __syncwarp();
    if (W::NVIS > 0 && need_visit) {
      // the tile's 32 visit layers: coalesced read-modify-write by the warp, crops captured on the way
      const int64_t e0 = tile * 32;
      const uint32_t cells = (uint32_t)(((p.n - e0) < 32 ? (p.n - e0) : 32) * (W::G * W::G));
      float *vis = p.visit + e0 * (W::G * W::G);
      float *cr = s_crop[W::NVIS > 0 ? warp : 0];
      constexpr uint32_t UN = 12;                                // 324 = 27 x 12 loads per lane, 12 in flight
      for (uint32_t k0 = 0; k0 < (uint32_t)(W::G * W::G); k0 += UN) {
        float vv[UN];
#pragma unroll
        for (uint32_t j = 0; j < UN; ++j) {
          const uint32_t idx = (k0 + j) * 32 + lane;
          vv[j] = (idx < cells && ((s_info[warp][idx / (W::G * W::G)] >> 20) & 3u) != 2u) ? __ldcs(vis + idx) : 0.0f;
        }
#pragma unroll
        for (uint32_t j = 0; j < UN; ++j) {
          const uint32_t idx = (k0 + j) * 32 + lane;
          if (idx >= cells) continue;
          const uint32_t env = idx / (W::G * W::G), cell = idx - env * (W::G * W::G);
          const int x = cell / W::G, y = cell - x * W::G;
          const uint32_t info = s_info[warp][env];
          const int bx = info & 31, by = (info >> 5) & 31, px = (info >> 10) & 31, py = (info >> 15) & 31;
          const uint32_t op = (info >> 20) & 3u;
          const int dx = x - bx + 2, dy = y - by + 2, qx = x - px + 2, qy = y - py + 2;
          const bool in_cur = dx >= 0 && dx < 5 && dy >= 0 && dy < 5;
          float val = vv[j];
          if (op == 1) val = visit_average(val, in_cur);
          else if (op == 2) val = W::visit_reset(in_cur);
          if (op) __stcs(vis + idx, val);
          if (in_cur) cr[env * 50 + dx * 5 + dy] = val;
          if (qx >= 0 && qx < 5 && qy >= 0 && qy < 5) cr[env * 50 + 25 + qx * 5 + qy] = val;
        }
      }
      __syncwarp();
    }
    // value of plane `slot`, cell `cell` of env `env` of this tile
    auto value = [&](uint32_t env, uint32_t slot, uint32_t cell) -> float {
      if (W::NVIS > 0 && (slot == (uint32_t)W::VIS_SLOT0 || slot == (uint32_t)W::VIS_SLOT1))
        return s_crop[W::NVIS > 0 ? warp : 0][env * 50 + (slot == (uint32_t)W::VIS_SLOT1 ? 25 : 0) + cell];
      return ((mk[env * W::NSLOT + slot] >> cell) & 1u) ? 1.0f : 0.0f;
    };
    // Lane l owns floats l, l+32, l+64, ... of EVERY env row, so which plane / cell a lane reads is fixed for the
    // whole kernel (no per-float index arithmetic) and a warp store covers 32 consecutive floats of one env.
    if (ff) {
      constexpr uint32_t PER = W::C * 25;                        // floats per env: the obs channels are slots 0..C-1
      constexpr int KK = (PER + 31) / 32;
      float *dst = reinterpret_cast<float *>(p.obs) + (tile * 32 - p.win_lo) * (int64_t)PER;
      for (unsigned m = ff; m; m &= m - 1) {
        const uint32_t env = __ffs(m) - 1;
#pragma unroll
        for (int k = 0; k < KK; ++k) {
          const uint32_t pos = lane + 32 * k;
          if (pos < PER) __stcs(dst + env * PER + pos, value(env, pos / 25, pos % 25));
        }
      }
    }
    if (W::HAS_LOC && fl) {
      constexpr uint32_t PER = 4 * 25;                           // local obs planes: slots 0, 7, 8, 3 (lmaze_env_v5.py:360-368)
      constexpr int KK = (PER + 31) / 32;
      float *dst = reinterpret_cast<float *>(p.obs2) + (tile * 32 - p.win_lo) * (int64_t)PER;
      for (unsigned m = fl; m; m &= m - 1) {
        const uint32_t env = __ffs(m) - 1;
        const bool err = (s_info[warp][env] >> 22) & 1u;         // IndexError in the reference: the row is all zero
#pragma unroll
        for (int k = 0; k < KK; ++k) {
          const uint32_t pos = lane + 32 * k, c = pos / 25;
          const uint32_t slot = c == 0 ? 9u : c == 1 ? 7u : c == 2 ? 8u : 10u;
          if (pos < PER) __stcs(dst + env * PER + pos, err ? 0.0f : value(env, slot, pos % 25));
        }
      }
    }
    tile = ntile; ntile = nntile;
  }
  if (p.mode == MODE_STEP) ws.flush(p.stats, lane);
  if (lane == 0) finish_grabber(p.work, (unsigned long long)gridDim.x * WARPS);
}

}  // namespace lmz
