/* TEST INFRASTRUCTURE ONLY (oracle/). Thin extern "C" harness around the
 * UNMODIFIED reference game core, compiled in place from
 * /root/reference/python/spacefortress/src/{config,configs,vector,object,
 * hexagon,game,wireframe}.cpp by oracle/Makefile into oracle/_ref/libsfref.so.
 * Nothing from the reference is copied into this repository; this file only
 * calls the reference's public C++ interface (game.hh:84-143, configs.hh:3-6).
 *
 * rand(): the reference calls the process-global libc rand() and never seeds
 * it (game.cpp:137-138,148). This file defines rand() itself and the Makefile
 * links with -Bsymbolic (and a private static libstdc++), so every Game instance gets its own glibc random_r() stream
 * (real glibc TYPE_3 generator, so this also pins the restatement in
 * sf_oracle.c against the real libc).
 *
 * Uninitialised members: Game::Game does `mFortress.mVulnerabilityTimer += 250`
 * on a member no constructor sets (game.cpp:78, game.cpp:12-14). The harness
 * constructs every Game with placement-new on zeroed memory, i.e. the
 * canonical reading "starts at 0 -> 250" (SURVEY.md §7.2, quirk Q5). */
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>

#include "space-fortress.hh"
#include "sf_record.h"

struct RefEnv {
  Config* config;
  Game* game;     /* placement-new'ed into mem */
  void* mem;
  struct random_data rd;
  char rstate[128];
  uint32_t seed;
  uint32_t count;
  int prev_vlner; /* ssf_env.py:92 */
  int shaped;     /* gametype in ["autoturn","youturn"], ssf_env.py:235 */
  int youturn;
  std::string gametype;
};

static RefEnv* g_current = nullptr;

/* Overrides libc rand() for the reference objects in this .so only: the link
 * step uses -Wl,-Bsymbolic-functions so game.o's call binds here. */
extern "C" int rand(void) noexcept {
  int32_t r = 0;
  random_r(&g_current->rd, &r);
  g_current->count++;
  return (int)r;
}

static Config* make_config(const std::string& name) {
  /* same dispatch as pymodule.cpp:331-343 */
  if (name == "autoturn") return autoturnConfig();
  if (name == "youturn") return youturnConfig();
  if (name == "test-youturn") return testyouturnConfig();
  if (name == "test-autoturn") return testautoturnConfig();
  return nullptr;
}

static void new_game(RefEnv* e) {
  /* ssf_env.py:164 creates a brand-new Game (and Config) on every reset */
  g_current = e;
  if (e->game) { e->game->~Game(); free(e->mem); }
  if (e->config) delete e->config;
  e->config = make_config(e->gametype);
  e->mem = calloc(1, sizeof(Game));
  e->game = new (e->mem) Game(e->config);
}

extern "C" {

RefEnv* sfref_create(const char* gametype, uint32_t seed) {
  std::string name(gametype);
  Config* probe = make_config(name);
  if (!probe) return nullptr;
  delete probe;
  RefEnv* e = new RefEnv();
  e->config = nullptr; e->game = nullptr; e->mem = nullptr;
  e->gametype = name;
  e->seed = seed; e->count = 0; e->prev_vlner = 0;
  e->shaped = (name == "autoturn" || name == "youturn");
  e->youturn = (name == "youturn" || name == "test-youturn");
  memset(&e->rd, 0, sizeof(e->rd));
  memset(e->rstate, 0, sizeof(e->rstate));
  initstate_r(seed, e->rstate, sizeof(e->rstate), &e->rd);
  new_game(e);
  return e;
}

void sfref_destroy(RefEnv* e) {
  if (!e) return;
  if (e->game) { e->game->~Game(); free(e->mem); }
  if (e->config) delete e->config;
  delete e;
}

/* SSF_Env.reset(): new Game; prev_vlner is NOT reset (ssf_env.py:163-178, :92) */
void sfref_reset(RefEnv* e) { new_game(e); }

int sfref_sizeof_game(void) { return (int)sizeof(Game); }

/* Raw core tick with a key mask, following the key-event order of
 * ssf_env.py:213-229. Returns the truncated int reward of pymodule.cpp:230. */
int sfref_core_step(RefEnv* e, int keymask) {
  g_current = e;
  Game* g = e->game;
  if (keymask & SFK_FIRE) g->pressKey(FIRE_KEY); else g->releaseKey(FIRE_KEY);
  if (keymask & SFK_THRUST) g->pressKey(THRUST_KEY); else g->releaseKey(THRUST_KEY);
  if (e->youturn) {
    if (keymask & SFK_LEFT) g->pressKey(LEFT_KEY); else g->releaseKey(LEFT_KEY);
    if (keymask & SFK_RIGHT) g->pressKey(RIGHT_KEY); else g->releaseKey(RIGHT_KEY);
  }
  return g->stepOneTick(34); /* tickdur, ssf_env.py:61 */
}

/* SSF_Env.step semantics (ssf_env.py:208-253) on top of the reference core:
 * out[0]=shaped reward, out[1]=done, out[2]=fort_kill, out[3]=event bits */
void sfref_env_step(RefEnv* e, int keymask, int* out) {
  Game* g = e->game;
  /* press edges and missed shots have no event string in the reference; derive them
   * from the public flags / counters around the tick */
  bool f0 = g->mShip.mFireFlag, t0 = g->mShip.mThrustFlag, l0 = g->mShip.mLeftFlag, r0 = g->mShip.mRightFlag;
  int missed0 = g->mStats.missedShots;
  int reward = sfref_core_step(e, keymask);
  int fort_kill = reward > 0;
  if (e->shaped) {
    int vl = g->mScore.mVulnerability;
    int change = vl - e->prev_vlner;
    if (vl <= 10 && !fort_kill) reward += change;
    if (reward > 1) reward = 1; else if (reward < -1) reward = -1;
    reward += 2 * fort_kill;
    e->prev_vlner = vl;
  }
  uint32_t ev = 0;
  for (size_t i = 0; i < g->mEvents.size(); i++) {
    const std::string& s = g->mEvents[i];
    if (s == "missile-fired") ev |= SFE_MISSILE_FIRED;
    else if (s == "fortress-fired") ev |= SFE_FORTRESS_FIRED;
    else if (s == "hit-fortress") ev |= SFE_HIT_FORTRESS;
    else if (s == "vlner-increased") ev |= SFE_VLNER_INCREASED;
    else if (s == "vlner-reset") ev |= SFE_VLNER_RESET;
    else if (s == "fortress-destroyed") ev |= SFE_FORTRESS_DESTROYED;
    else if (s == "hit-dead-fortress") ev |= SFE_HIT_DEAD_FORTRESS;
    else if (s == "explode-bighex") ev |= SFE_EXPLODE_BIGHEX;
    else if (s == "explode-smallhex") ev |= SFE_EXPLODE_SMALLHEX;
    else if (s == "shell-hit-ship") ev |= SFE_SHELL_HIT_SHIP;
    else if (s == "ship-respawn") ev |= SFE_SHIP_RESPAWN;
    else if (s == "fortress-respawn") ev |= SFE_FORTRESS_RESPAWN;
  }
  if (!f0 && g->mShip.mFireFlag) ev |= SFE_PRESS_FIRE;
  if (!t0 && g->mShip.mThrustFlag) ev |= SFE_PRESS_THRUST;
  if (!l0 && g->mShip.mLeftFlag) ev |= SFE_PRESS_LEFT;
  if (!r0 && g->mShip.mRightFlag) ev |= SFE_PRESS_RIGHT;
  if (g->mStats.missedShots != missed0) ev |= SFE_MISSED_SHOT;
  if (g->mCollisions.bigHex) ev |= SFE_COL_BIGHEX;
  if (g->mCollisions.smallHex) ev |= SFE_COL_SMALLHEX;
  if (g->mCollisions.missileFortress) ev |= SFE_COL_MISSILE_FORTRESS;
  if (g->mCollisions.shellShip) ev |= SFE_COL_SHELL_SHIP;
  out[0] = reward;
  out[1] = g->isGameOver() ? 1 : 0;
  out[2] = fort_kill;
  out[3] = (int)ev;
}

void sfref_get_state(RefEnv* e, sfr_record* r) {
  Game* g = e->game;
  memset(r, 0, sizeof(*r));
  r->ship_x = g->mShip.mPos.mX; r->ship_y = g->mShip.mPos.mY;
  r->ship_vx = g->mShip.mVel.mX; r->ship_vy = g->mShip.mVel.mY;
  r->ship_angle = g->mShip.mAngle;
  r->fortress_angle = g->mFortress.mAngle;
  r->fortress_last_angle = g->mFortress.mLastAngle;
  for (int i = 0; i < MAX_MISSILES; i++) {
    const Object& o = g->mMissiles[i];
    if (o.mAlive) {
      r->missile_mask |= 1u << i;
      r->missile_x[i] = o.mPos.mX; r->missile_y[i] = o.mPos.mY;
      r->missile_vx[i] = o.mVel.mX; r->missile_vy[i] = o.mVel.mY;
      r->missile_angle[i] = o.mAngle;
    }
  }
  for (int i = 0; i < MAX_SHELLS; i++) {
    const Object& o = g->mShells[i];
    if (o.mAlive) {
      r->shell_mask |= 1u << i;
      r->shell_x[i] = o.mPos.mX; r->shell_y[i] = o.mPos.mY;
      r->shell_vx[i] = o.mVel.mX; r->shell_vy[i] = o.mVel.mY;
      r->shell_angle[i] = o.mAngle;
    }
  }
  r->points = g->mScore.mPoints; r->raw_points = g->mScore.mRawPoints;
  r->ship_alive = g->mShip.mAlive; r->fortress_alive = g->mFortress.mAlive;
  r->ship_death_timer = g->mShip.mDeathTimer;
  r->fire_timer = g->mShip.mFireTimer; r->thrust_timer = g->mShip.mThrustTimer;
  r->left_timer = g->mShip.mLeftTimer; r->right_timer = g->mShip.mRightTimer;
  r->thrust_flag = g->mShip.mThrustFlag; r->fire_flag = g->mShip.mFireFlag;
  r->left_flag = g->mShip.mLeftFlag; r->right_flag = g->mShip.mRightFlag;
  r->turn_flag = (int)g->mShip.mTurnFlag;
  r->fortress_timer = g->mFortress.mTimer;
  r->fortress_death_timer = g->mFortress.mDeathTimer;
  r->fortress_vuln_timer = g->mFortress.mVulnerabilityTimer;
  r->vulnerability = g->mScore.mVulnerability;
  r->tick = g->mTick; r->time = g->mTime;
  const Stats& s = g->mStats;
  int st[SFR_NUM_STATS] = {s.bigHexDeaths, s.smallHexDeaths, s.shellDeaths, s.shipDeaths,
                           s.resets, s.destroyedFortresses, s.missedShots, s.totalShots,
                           s.totalThrusts, s.totalLefts, s.totalRights, s.vlnerIncs, s.maxVlner};
  memcpy(r->stats, st, sizeof(st));
  r->prev_vlner = e->prev_vlner;
  r->rng_seed = e->seed; r->rng_count = e->count;
}

/* Teacher forcing: overwrite the public members of the live Game with a record
 * (all members are public, game.hh:85-107). The rand() stream is re-derived
 * from (seed, count). */
void sfref_set_state(RefEnv* e, const sfr_record* r) {
  Game* g = e->game;
  g->mShip.mPos.mX = r->ship_x; g->mShip.mPos.mY = r->ship_y;
  g->mShip.mVel.mX = r->ship_vx; g->mShip.mVel.mY = r->ship_vy;
  g->mShip.mAngle = r->ship_angle;
  g->mFortress.mAngle = r->fortress_angle;
  g->mFortress.mLastAngle = r->fortress_last_angle;
  for (int i = 0; i < MAX_MISSILES; i++) {
    Object& o = g->mMissiles[i];
    o.mAlive = (r->missile_mask >> i) & 1;
    o.mPos.mX = r->missile_x[i]; o.mPos.mY = r->missile_y[i];
    o.mVel.mX = r->missile_vx[i]; o.mVel.mY = r->missile_vy[i];
    o.mAngle = r->missile_angle[i];
    o.mCollisionRadius = 5;
  }
  for (int i = 0; i < MAX_SHELLS; i++) {
    Object& o = g->mShells[i];
    o.mAlive = (r->shell_mask >> i) & 1;
    o.mPos.mX = r->shell_x[i]; o.mPos.mY = r->shell_y[i];
    o.mVel.mX = r->shell_vx[i]; o.mVel.mY = r->shell_vy[i];
    o.mAngle = r->shell_angle[i];
    o.mCollisionRadius = 3;
  }
  g->mScore.mPoints = r->points; g->mScore.mRawPoints = r->raw_points;
  g->mShip.mAlive = r->ship_alive != 0; g->mFortress.mAlive = r->fortress_alive != 0;
  g->mShip.mDeathTimer = r->ship_death_timer;
  g->mShip.mFireTimer = r->fire_timer; g->mShip.mThrustTimer = r->thrust_timer;
  g->mShip.mLeftTimer = r->left_timer; g->mShip.mRightTimer = r->right_timer;
  g->mShip.mThrustFlag = r->thrust_flag != 0; g->mShip.mFireFlag = r->fire_flag != 0;
  g->mShip.mLeftFlag = r->left_flag != 0; g->mShip.mRightFlag = r->right_flag != 0;
  g->mShip.mTurnFlag = (Turn)r->turn_flag;
  g->mFortress.mTimer = r->fortress_timer;
  g->mFortress.mDeathTimer = r->fortress_death_timer;
  g->mFortress.mVulnerabilityTimer = r->fortress_vuln_timer;
  g->mScore.mVulnerability = r->vulnerability;
  g->mTick = r->tick; g->mTime = r->time;
  Stats& s = g->mStats;
  s.bigHexDeaths = r->stats[0]; s.smallHexDeaths = r->stats[1]; s.shellDeaths = r->stats[2];
  s.shipDeaths = r->stats[3]; s.resets = r->stats[4]; s.destroyedFortresses = r->stats[5];
  s.missedShots = r->stats[6]; s.totalShots = r->stats[7]; s.totalThrusts = r->stats[8];
  s.totalLefts = r->stats[9]; s.totalRights = r->stats[10]; s.vlnerIncs = r->stats[11];
  s.maxVlner = r->stats[12];
  e->prev_vlner = r->prev_vlner;
  e->seed = r->rng_seed;
  memset(&e->rd, 0, sizeof(e->rd));
  memset(e->rstate, 0, sizeof(e->rstate));
  initstate_r(e->seed, e->rstate, sizeof(e->rstate), &e->rd);
  e->count = 0;
  g_current = e;
  for (uint32_t i = 0; i < r->rng_count; i++) rand();
}

/* Game::dumpState(), game.cpp:519-576 */
int sfref_dump(RefEnv* e, char* buf, int cap) {
  std::string s = e->game->dumpState();
  int n = (int)s.size();
  if (n >= cap) n = cap - 1;
  memcpy(buf, s.data(), n);
  buf[n] = 0;
  return n;
}

/* feature-obs extras, game.cpp:282-312 (mExtra is public) */
void sfref_get_extra(RefEnv* e, double* out4) {
  out4[0] = e->game->mExtra.vdir; out4[1] = e->game->mExtra.fdist;
  out4[2] = e->game->mExtra.ndist; out4[3] = e->game->mExtra.aim;
}

/* Game.events (pymodule.cpp:136-143): the tick's event strings in order, joined with ','; and Game.collisions
 * (pymodule.cpp:182-197), which fills its tuple from the back: shell, missile, smallhex, bighex */
int sfref_events(RefEnv* e, char* buf, int cap) {
  std::string s;
  for (size_t i = 0; i < e->game->mEvents.size(); i++) { if (i) s += ","; s += e->game->mEvents[i]; }
  int n = (int)s.size();
  if (n >= cap) n = cap - 1;
  memcpy(buf, s.data(), n);
  buf[n] = 0;
  return n;
}
int sfref_collisions(RefEnv* e, char* buf, int cap) {
  const Collisions& c = e->game->mCollisions;
  std::string s;
  if (c.shellShip) s += "shell,";
  if (c.missileFortress) s += "missile,";
  if (c.smallHex) s += "smallhex,";
  if (c.bigHex) s += "bighex,";
  if (!s.empty()) s.pop_back();
  int n = (int)s.size();
  if (n >= cap) n = cap - 1;
  memcpy(buf, s.data(), n);
  buf[n] = 0;
  return n;
}

/* hexagon vertices (hexagon.cpp:13-35): out[12] = big x0,y0..x5,y5 ; small likewise */
void sfref_hexagons(RefEnv* e, double* big12, double* small12) {
  for (int i = 0; i < 6; i++) {
    big12[2 * i] = e->game->mBighex.mPoints[i].mX; big12[2 * i + 1] = e->game->mBighex.mPoints[i].mY;
    small12[2 * i] = e->game->mSmallhex.mPoints[i].mX; small12[2 * i + 1] = e->game->mSmallhex.mPoints[i].mY;
  }
}

/* wireframe tables (wireframe.cpp:8-70): which: 0 missile 1 shell 2 ship 3 fortress
 * out: n_lines, then per line x0,y0,x1,y1 (as doubles) */
int sfref_wireframe(int which, double* out, int cap) {
  initWireframes();
  WireFrame* wf = which == 0 ? &missileWireFrame : which == 1 ? &shellWireFrame
                : which == 2 ? &shipWireFrame : &fortressWireFrame;
  int n = wf->lineCount;
  if (4 * n > cap) return -1;
  for (int i = 0; i < n; i++) {
    out[4 * i + 0] = wf->points[wf->lines[i].from].mX; out[4 * i + 1] = wf->points[wf->lines[i].from].mY;
    out[4 * i + 2] = wf->points[wf->lines[i].to].mX; out[4 * i + 3] = wf->points[wf->lines[i].to].mY;
  }
  return n;
}

/* Bulk throughput helper for bench.py's cpu_baseline/reference arm: run `steps`
 * env steps (auto-reset on done, like the gym_vecenv worker loop) with the
 * given key masks; returns the sum of shaped rewards (so the work cannot be
 * optimised away). No rendering here — the renderer cannot be built (no cairo). */
long sfref_run(RefEnv* e, const unsigned char* keymasks, long steps) {
  long acc = 0;
  int out[4];
  for (long i = 0; i < steps; i++) {
    sfref_env_step(e, keymasks[i], out);
    acc += out[0];
    if (out[1]) sfref_reset(e);
  }
  return acc;
}

/* Same loop with a frame per step: the reference draws with cairo (draw.cpp), which cannot be built
 * offline, so the caller passes the restated renderer (oracle/sf_draw_oracle.c: sfo_draw_obs) as a
 * function pointer; the game tick, shaping and auto-reset are the reference's own code. */
typedef void (*sfref_draw_fn)(const sfr_record*, unsigned char*);
long sfref_run_render(RefEnv* e, const unsigned char* keymasks, long steps, sfref_draw_fn draw, unsigned char* obs84) {
  long acc = 0;
  int out[4];
  sfr_record rec;
  for (long i = 0; i < steps; i++) {
    sfref_env_step(e, keymasks[i], out);
    acc += out[0];
    if (out[1]) sfref_reset(e);
    sfref_get_state(e, &rec);
    draw(&rec, obs84);
    acc += obs84[(i * 7) % (84 * 84)];
  }
  return acc;
}

} /* extern "C" */
