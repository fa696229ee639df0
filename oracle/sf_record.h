/* TEST INFRASTRUCTURE ONLY (oracle/): per-env game-state record used to move
 * state between the compiled reference core (oracle/_ref), the C restatement
 * (oracle/sf_oracle.c) and the tests. Mirrors the public members of the
 * reference's Game object (python/spacefortress/src/game.hh:84-107) plus the
 * one piece of Python-layer state (prev_vlner, ssf_env.py:92,244) and the
 * libc rand() stream position (game.cpp:137-148).
 *
 * The product's C-ABI record (include/sf_b200.h: sf_state_record) has the same
 * layout on purpose; tests assert sizeof equality and compare field by field. */
#ifndef SF_RECORD_H
#define SF_RECORD_H
#include <stdint.h>

#define SFR_MAX_MISSILES 20
#define SFR_MAX_SHELLS 20
#define SFR_NUM_STATS 13

typedef struct sfr_record {
  double ship_x, ship_y, ship_vx, ship_vy, ship_angle;
  double fortress_angle, fortress_last_angle;
  double missile_x[SFR_MAX_MISSILES], missile_y[SFR_MAX_MISSILES];
  double missile_vx[SFR_MAX_MISSILES], missile_vy[SFR_MAX_MISSILES];
  double missile_angle[SFR_MAX_MISSILES];
  double shell_x[SFR_MAX_SHELLS], shell_y[SFR_MAX_SHELLS];
  double shell_vx[SFR_MAX_SHELLS], shell_vy[SFR_MAX_SHELLS];
  double shell_angle[SFR_MAX_SHELLS];
  float points, raw_points;
  uint32_t missile_mask, shell_mask; /* bit i = slot i alive */
  int32_t ship_alive, fortress_alive;
  int32_t ship_death_timer, fire_timer, thrust_timer, left_timer, right_timer;
  int32_t thrust_flag, fire_flag, left_flag, right_flag, turn_flag;
  int32_t fortress_timer, fortress_death_timer, fortress_vuln_timer;
  int32_t vulnerability, tick, time;
  int32_t stats[SFR_NUM_STATS]; /* order of game.hh:29-43 */
  int32_t prev_vlner;           /* ssf_env.py:92,244 */
  uint32_t rng_seed;            /* srand() seed of this env's stream */
  uint32_t rng_count;           /* rand() calls consumed so far */
  int32_t _pad;
} sfr_record;

/* per-step event bits (replaces the string list of game.cpp:124-127) */
enum {
  SFE_MISSILE_FIRED = 1 << 0,
  SFE_FORTRESS_FIRED = 1 << 1,
  SFE_HIT_FORTRESS = 1 << 2,
  SFE_VLNER_INCREASED = 1 << 3,
  SFE_VLNER_RESET = 1 << 4,
  SFE_FORTRESS_DESTROYED = 1 << 5,
  SFE_HIT_DEAD_FORTRESS = 1 << 6,
  SFE_EXPLODE_BIGHEX = 1 << 7,
  SFE_EXPLODE_SMALLHEX = 1 << 8,
  SFE_SHELL_HIT_SHIP = 1 << 9,
  SFE_SHIP_RESPAWN = 1 << 10,
  SFE_FORTRESS_RESPAWN = 1 << 11,
  /* Collisions struct, game.hh:45-47 */
  SFE_COL_BIGHEX = 1 << 12,
  SFE_COL_SMALLHEX = 1 << 13,
  SFE_COL_MISSILE_FORTRESS = 1 << 14,
  SFE_COL_SHELL_SHIP = 1 << 15,
  /* key edges (press-/release- events, game.cpp:223) */
  SFE_PRESS_FIRE = 1 << 16,
  SFE_PRESS_THRUST = 1 << 17,
  SFE_PRESS_LEFT = 1 << 18,
  SFE_PRESS_RIGHT = 1 << 19,
  SFE_MISSED_SHOT = 1 << 20,
  SFE_EPISODE_RESET = 1 << 21 /* auto-reset happened after this step */
};

/* key mask bits: the env sends FIRE, THRUST[, LEFT, RIGHT] every step
 * (ssf_env.py:213-229) */
enum { SFK_FIRE = 1, SFK_THRUST = 2, SFK_LEFT = 4, SFK_RIGHT = 8 };

#endif
