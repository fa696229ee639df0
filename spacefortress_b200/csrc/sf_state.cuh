// sf_state.cuh — SoA device state of the batched simulator and its per-env load/store helpers.
//
// One env = one index into every array; all per-step scalars are grouped into 128-bit vectors so a warp
// loads 512 contiguous bytes per instruction. Replaces the reference's heap-allocated AoS `Game`
// (game.hh:84-107, sizeof 2568 B, one `new Game` per episode: ssf_env.py:164).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "sf_tables.h"

#define SF_DEV_SHELLS 4  // <= 3 shells can be alive (shell life <= 80 ticks, fire period >= 30 ticks)
#define SF_RNG_WORDS 31  // glibc TYPE_3 ring (game.cpp:137-148 call rand())

// q0.x "core" word
#define SF_CORE_ANGLE_MASK 0x1FFu
#define SF_CORE_FANG_SHIFT 9    // fortress mAngle / 10        (6 bits)
#define SF_CORE_FLAST_SHIFT 15  // fortress mLastAngle / 10    (6 bits)
#define SF_CORE_SHIP_ALIVE (1u << 21)
#define SF_CORE_FORT_ALIVE (1u << 22)
#define SF_CORE_FIRE (1u << 23)
#define SF_CORE_THRUST (1u << 24)
#define SF_CORE_LEFT (1u << 25)
#define SF_CORE_RIGHT (1u << 26)
// q0.y "pmask": bits 0..19 missile slots alive, bits 20..23 shell slots alive
#define SF_PMASK_MISSILES 0xFFFFFu
#define SF_PMASK_SHELL_SHIFT 20

struct SfDev {
  int n;         // envs
  int n_pad;     // array pitch (multiple of 32)
  // ship kinematics (game.hh:58-73 via object.hh:3-14): fp64 like the reference
  double2* pos;  // [n] x,y
  double2* vel;  // [n] vx,vy
  int4* q0;      // core, pmask, ship.mDeathTimer, fortress.mTimer
  int4* q1;      // fortress.mDeathTimer, fortress.mVulnerabilityTimer, mVulnerability, prev_vlner (ssf_env.py:92)
  int4* q2;      // fire/thrust/left/right key timers (game.hh:61-64)
  int4* q3;      // mPoints(f32 bits), mRawPoints(f32 bits), mTick, episode return
  int4* st0;     // bigHexDeaths smallHexDeaths shellDeaths shipDeaths      (game.hh:29-43)
  int4* st1;     // resets destroyedFortresses missedShots totalShots
  int4* st2;     // totalThrusts totalLefts totalRights vlnerIncs
  int4* st3;     // maxVlner, rng ring index, rng calls consumed, rng seed
  double2* mpos; // [20][n_pad] missile x,y (velocity is 20*(cos,sin)(angle): not stored)
  short* mang;   // [20][n_pad] missile angle, integer degrees
  double2* spos; // [4][n_pad]
  double2* svel; // [4][n_pad]
  double* sang;  // [4][n_pad] shell angle (real valued, game.cpp:166)
  unsigned* rng; // [31][n_pad]
  unsigned char* expc;        // [n][28*28] ship-explosion sprite cache (render memo; not game state)
  unsigned* expstamp;         // [n] which life's explosion the cache holds: rand() calls consumed at its spawn + 1 (0: none)
  unsigned char* expo;        // [n][32*32] RESAMPLED explosion box (output pixels, row stride 32) without any wireframe layer
  uint2* expo_meta;           // [n] {life, fortress state | bar state<<6 | (points & 0x3FFFF)<<10 | valid quarters<<28}: what expo holds
  unsigned long long* epi;    // [SF_NUM_EPISODE_STATS] finished-episode accumulators
  const SfTables* tab;
  const unsigned char* static_image;  // the static part of the rendering blocks' shared memory, ready to be bulk-copied (sf_pack_static_kernel)
  // preset (configs.cpp:51-89)
  int autoturn, shaped, destroy_fortress, death_penalty;
  float missile_penalty;
  int num_actions;
  int keymask_of_action[16];
  long long first_global_env;
};
