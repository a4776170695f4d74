// lmz_variants.h -- compile-time description of each maze variant and of the
// shared-memory "template blob" its observations are assembled from.
//
// An observation is never computed pixel by pixel.  For a fixed maze every
// channel of the reference's upsampled image is either static (walls / goal /
// blanks, lmaze_env.py:92-107) or a single ExE block of ones on a zero plane
// (the ball, lmaze_env.py:80; v3 also the goal one-hot, lmaze_env_v3.py:167,182).
// So each env's obs is the concatenation of a handful of SEGMENTS of one
// per-variant byte blob:
//
//   v0 blob:  [ZERO 77 rows][STATIC ch1..ch3][BAND table: 12 x (7 rows with a 7-wide run of 1.0 at column 7y)]
//     obs(x,y) = ZERO[0 : 7x rows] | BAND[y] | ZERO[7x rows : end] + STATIC      (3 segments)
//   v3 blob:  [STATIC ch0][ZERO 136 rows][BAND table: 18 x (4 rows, run at 4y)]
//     obs      = STATIC + ZERO[0 : 4x] | BAND[y] | ZERO[0 : (68-4x)+4gx] | BAND[gy] | ZERO[0 : 68-4gx]   (5 segments)
//
// All segment offsets and lengths are multiples of 16 bytes (row = 336 B / 288 B),
// which is what both cp.async.bulk and st.global.v4 need.  The blob also carries
// the G*G cell-class table and the spawn-candidate table, so one bulk load stages
// everything a CTA needs.  The blob ends with the un-expanded static layers twice more: as bytes (u8 [C][G][G], the
// compact observation) and as bits (bit k of word k/32 = cell k of the flattened [C][G][G] layers, the bit-packed
// observation); an env's compact / bit-packed row is that template with its one-hot cell(s) set.
#pragma once
#include <stdint.h>

namespace lmz {

enum : int { CLS_W = 0, CLS_B = 1, CLS_X = 2, CLS_S = 3 };

// Reward codes kept in the packed state (v0 needs the LAST reward: moving into
// the 'S' cell takes no branch and returns it again, lmaze_env.py:172-196).
enum : int { RC_NEG_ZERO = 0, RC_WALL = 1, RC_MOVE = 2, RC_GOAL = 3 };

struct Seg {            // one contiguous piece of an env's observation
  uint32_t dst;         // byte offset inside the env's obs
  uint32_t src;         // byte offset inside the blob
  uint32_t len;         // bytes (multiple of 16, may be 0)
};

constexpr uint32_t align16(uint32_t v) { return (v + 15u) & ~15u; }

struct EnvRegs {        // unpacked per-env state as held in registers
  int x, y, gx, gy;
  uint32_t step;
  int rcode;
};

// ---------------------------------------------------------------- v0
struct V0 {
  static constexpr int ID = 0;
  static constexpr int G = 12, E = 7, C = 4, S = G * E;              // lmaze_env.py:17-20
  static constexpr uint32_t ROW_BYTES = S * 4;                       // 336
  static constexpr uint32_t PLANE_BYTES = ROW_BYTES * S;             // 28,224
  static constexpr uint32_t OBS_BYTES = PLANE_BYTES * C;             // 112,896
  static constexpr uint32_t BAND_BYTES = ROW_BYTES * E;              // 2,352
  static constexpr uint32_t ZERO_ROWS = S - E;                       // 77
  static constexpr uint32_t ZERO_OFF = 0;
  static constexpr uint32_t STATIC_OFF = ZERO_OFF + ZERO_ROWS * ROW_BYTES;
  static constexpr uint32_t BAND_OFF = STATIC_OFF + 3 * PLANE_BYTES;
  static constexpr uint32_t CLS_OFF = BAND_OFF + G * BAND_BYTES;
  static constexpr uint32_t CAND_OFF = CLS_OFF + align16(G * G);
  static constexpr int MAX_CAND = 80;
  static constexpr uint32_t COMPACT_OFF = CAND_OFF + align16(MAX_CAND * 2);   // u8 [C][G][G] un-expanded layers
  static constexpr uint32_t COMPACT_BYTES = C * G * G;                         // 576
  static constexpr uint32_t BITS_OFF = COMPACT_OFF + align16(COMPACT_BYTES);    // the same layers, one BIT per cell
  static constexpr uint32_t BITS_WORDS = (C * G * G + 31) / 32;                // 18 words = 72 B per env
  static constexpr uint32_t BLOB_BYTES = BITS_OFF + align16(BITS_WORDS * 4);
  static constexpr uint32_t TABLES_OFF = CLS_OFF, TABLES_BYTES = BLOB_BYTES - CLS_OFF;
  static constexpr int NHOT = 1;                                               // one-hot bytes per env in the compact obs
  static constexpr int NSEG = 3;
  static constexpr uint32_t STEP_SAT = (1u << 20) - 1;
  static constexpr int STEP_LIMIT = 100;                             // lmaze_env.py:247

  // packed state: x:5 | y:5 | rcode:2 | step:20 (saturating)
  __host__ __device__ static inline EnvRegs unpack(uint32_t s) {
    EnvRegs r;
    r.x = s & 31; r.y = (s >> 5) & 31; r.rcode = (s >> 10) & 3; r.step = s >> 12;
    r.gx = 5; r.gy = 5;
    return r;
  }
  __host__ __device__ static inline uint32_t pack(const EnvRegs &r) {
    return (uint32_t)r.x | ((uint32_t)r.y << 5) | ((uint32_t)r.rcode << 10) | (r.step << 12);
  }
  // compact obs = static u8 layers with the ball's byte set in layer 0 (lmaze_env.py:80)
  __host__ __device__ static inline void hot_bytes(const EnvRegs &r, uint32_t *hot) { hot[0] = (uint32_t)(r.x * G + r.y); }
  __host__ __device__ static inline void segments(const EnvRegs &r, Seg *sg) {
    const uint32_t top = (uint32_t)r.x * E * ROW_BYTES;               // rows above the ball band
    sg[0] = {0u, ZERO_OFF, top};
    sg[1] = {top, BAND_OFF + (uint32_t)r.y * BAND_BYTES, BAND_BYTES};
    sg[2] = {top + BAND_BYTES, ZERO_OFF + top, (ZERO_ROWS * ROW_BYTES - top) + 3 * PLANE_BYTES};
  }
};

// ---------------------------------------------------------------- v3
struct V3 {
  static constexpr int ID = 3;
  static constexpr int G = 18, E = 4, C = 3, S = G * E;              // lmaze_env_v3.py:72-76,87-89
  static constexpr uint32_t ROW_BYTES = S * 4;                       // 288
  static constexpr uint32_t PLANE_BYTES = ROW_BYTES * S;             // 20,736
  static constexpr uint32_t OBS_BYTES = PLANE_BYTES * C;             // 62,208
  static constexpr uint32_t BAND_BYTES = ROW_BYTES * E;              // 1,152
  static constexpr uint32_t ZERO_ROWS = 2 * (S - E);                 // 136
  static constexpr uint32_t STATIC_OFF = 0;
  static constexpr uint32_t ZERO_OFF = STATIC_OFF + PLANE_BYTES;
  static constexpr uint32_t BAND_OFF = ZERO_OFF + ZERO_ROWS * ROW_BYTES;
  static constexpr uint32_t CLS_OFF = BAND_OFF + G * BAND_BYTES;
  static constexpr uint32_t CAND_OFF = CLS_OFF + align16(G * G);
  static constexpr int MAX_CAND = 80;
  static constexpr uint32_t COMPACT_OFF = CAND_OFF + align16(MAX_CAND * 2);
  static constexpr uint32_t COMPACT_BYTES = C * G * G;                         // 972
  static constexpr uint32_t BITS_OFF = COMPACT_OFF + align16(COMPACT_BYTES);
  static constexpr uint32_t BITS_WORDS = (C * G * G + 31) / 32;                // 31 words = 124 B per env (972 bits + 20 pad)
  static constexpr uint32_t BLOB_BYTES = BITS_OFF + align16(BITS_WORDS * 4);
  static constexpr uint32_t TABLES_OFF = CLS_OFF, TABLES_BYTES = BLOB_BYTES - CLS_OFF;
  static constexpr int NHOT = 2;
  static constexpr int NSEG = 5;
  static constexpr uint32_t STEP_SAT = (1u << 12) - 1;
  static constexpr int STEP_LIMIT = 100;                             // lmaze_env_v3.py:99

  // packed state: x:5 | y:5 | gx:5 | gy:5 | step:12 (saturating; done is step > 100)
  __host__ __device__ static inline EnvRegs unpack(uint32_t s) {
    EnvRegs r;
    r.x = s & 31; r.y = (s >> 5) & 31; r.gx = (s >> 10) & 31; r.gy = (s >> 15) & 31;
    r.step = s >> 20; r.rcode = 0;
    return r;
  }
  __host__ __device__ static inline uint32_t pack(const EnvRegs &r) {
    return (uint32_t)r.x | ((uint32_t)r.y << 5) | ((uint32_t)r.gx << 10) | ((uint32_t)r.gy << 15) | (r.step << 20);
  }
  // compact obs: layer 1 = ball one-hot, layer 2 = goal one-hot (lmaze_env_v3.py:167,182)
  __host__ __device__ static inline void hot_bytes(const EnvRegs &r, uint32_t *hot) {
    hot[0] = (uint32_t)(G * G + r.x * G + r.y); hot[1] = (uint32_t)(2 * G * G + r.gx * G + r.gy);
  }
  __host__ __device__ static inline void segments(const EnvRegs &r, Seg *sg) {
    const uint32_t btop = (uint32_t)r.x * E * ROW_BYTES;
    const uint32_t gtop = (uint32_t)r.gx * E * ROW_BYTES;
    const uint32_t tail = (S - E) * ROW_BYTES;                        // zero rows of a one-hot plane
    sg[0] = {0u, STATIC_OFF, PLANE_BYTES + btop};                    // ch0 + rows above the ball
    sg[1] = {PLANE_BYTES + btop, BAND_OFF + (uint32_t)r.y * BAND_BYTES, BAND_BYTES};
    sg[2] = {PLANE_BYTES + btop + BAND_BYTES, ZERO_OFF, (tail - btop) + gtop};
    sg[3] = {2 * PLANE_BYTES + gtop, BAND_OFF + (uint32_t)r.gy * BAND_BYTES, BAND_BYTES};
    sg[4] = {2 * PLANE_BYTES + gtop + BAND_BYTES, ZERO_OFF, tail - gtop};
  }
};

}  // namespace lmz
