// sf_tables.h — static tables built once on the host at sf_create() and uploaded to the GPU.
// Everything here depends only on constants of the reference (configs.cpp, wireframe.cpp, draw.cpp),
// never on per-env state: background hexagons, the fortress (fixed position, 36 possible angles),
// the fortress explosion, score digits, resize weights, trig LUTs for integer degrees.
#pragma once
#include <stdint.h>

#define SF_NAT_W 90  // int(450*.2), ssf_env.py:57
#define SF_NAT_H 92  // int(460*.2), ssf_env.py:58

#define SF_NAT_STRIDE 96           // row stride of the bg_nat table (16-byte aligned rows)
#define SF_TILE_STRIDE 92          // row stride of the per-warp shared-memory tile (4-byte aligned rows)
#define SF_FORT_W 17               // fortress sprite box (native px)
#define SF_FORT_X0 37
#define SF_FORT_Y0 39
#define SF_EXP_W 28                // explosion box: centre px - 13 .. centre px + 14
#define SF_EXPO_STRIDE 32           // resampled explosion box cache: 28 native columns / rows feed <= 28 output ones
#define SF_EXPO_BYTES (32 * SF_EXPO_STRIDE)
#define SF_EXP_LAYERS 4
#define SF_FEXP_X0 (45 - 13)
#define SF_FEXP_Y0 (47 - 13)
#define SF_TEXT_X0 32              // score strip (native px)
#define SF_TEXT_Y0 1
#define SF_TEXT_W 27
#define SF_TEXT_H 5
#define SF_BAR_X0 25
#define SF_BAR_Y0 88
#define SF_BAR_W 40
#define SF_BAR_H 3
#define SF_EXP_STROKES 85          // 7 rings x 12 arcs + the r=7 circle
#define SF_EXP_QUADS 100           // 84 arc quads + 16 circle quads
#define SF_MAX_TAPS 3
#define SF_MAGIC_N 512              // magic reciprocals floor((2^32-1)/d)+1 for edge heights d < 512 sub-rows
#define SF_FORT_LIST 160           // max lit pixels of a fortress sprite
#define SF_OBS_CHUNKS 441          // 84*84/16
#define SF_BAR_CHUNK0 420          // output rows 80..83 == chunks 420..440
#define SF_BAR_STATES 12           // filled 0..10 steps in grey .66, or full white (kill window)
#define SF_FORT_STATES 37          // 36 sector angles + destroyed (explosion)
#define SF_FORT_CHUNKS 160         // upper bound of output chunks influenced by the fortress/explosion box

// Ship explosion, scan-converted ahead of time. The 100 quads of drawExplosion (draw.cpp:116-145) are fixed
// relative to the 24.8 fixed-point centre c. In x the scan conversion is translation invariant; in y the rounding of
// the corners to the 1/15 px sample grid depends only on c.y & 255. So for each of the 256 y phases the exact span
// [lo, hi) of every quad on every sample row is tabulated relative to c (x in 1/256 px, rows relative to the
// centre's pixel row), grouped into (quad, pixel row) items; the device only adds the spans of an item into its
// <= SF_EXPT_NC cells.
#define SF_EXPT_ITEMS 224
#define SF_EXPT_SPANS 1792
#define SF_EXPT_NC 6
struct alignas(8) SfExpItem { unsigned char quad; signed char row; unsigned char n, pad; unsigned short span0; unsigned short pad2; };  // n spans from span0, pixel row `row`
struct alignas(16) SfExpPhase {
  int n_items, n_spans, pad[2];
  SfExpItem item[SF_EXPT_ITEMS];
  short span[SF_EXPT_SPANS][2];           // lo, hi relative to c.x
  unsigned short item0[SF_EXP_QUADS];     // first item of a quad (its pixel rows are consecutive items)
  signed char row0[SF_EXP_QUADS];         // first pixel row of a quad relative to the centre's
  short qxmin[SF_EXP_QUADS];              // leftmost corner of a quad relative to c.x: its cells start at (c.x + qxmin) >> 8
};

// the few tables the game step reads every tick, together: one copy lives in SfTables (global memory: state-only
// kernel), one in the shared memory of every rendering block
struct SfHot {
  double cs[360][2];        // cos, sin of integer degrees (== cos_deg, sin_deg)
  double hex[2][6][4];      // per hexagon edge: px, py, nx, ny (== hex_px, hex_py, hex_nx, hex_ny)
  double atan2_oct[8];
  double ship_start_vx, ship_start_vy;
  // Object::collided (object.cpp:12-15) is sqrt(dx*dx + dy*dy) <= r. sqrt is correctly rounded and monotone, so the test
  // equals dx*dx + dy*dy <= T(r) with T(r) = the largest double whose rounded square root is <= r (found by stepping
  // through the doubles around r*r on the host, sf_tables.cpp): [0] r = 13 shell-ship, [1] r = 23 missile-fortress,
  // [2] r = 21 the shell hide radius of draw.cpp:249-250.
  double touch2[4];
};

struct SfTap { int si, cnt; float a[SF_MAX_TAPS]; };  // consecutive source indices si..si+cnt-1 and their weights

struct SfTables {
  // trig for integer degrees, computed with the host libm exactly as the reference does:
  // cos(deg2rad(a)) with deg2rad(a) = a*M_PI/180 (vector.cpp:34-36, game.cpp:184-185,328-329)
  double cos_deg[360], sin_deg[360];
  // atan2(dy,dx) for the 8 directions where a ceil() downstream could flip on a 1-ulp difference
  // (index = octant*45 degrees: E, SE(+y), S, SW, W, NW, N, NE in screen coordinates), from the host libm
  double atan2_oct[8];
  double ship_start_vx, ship_start_vy;  // cos/sin(deg2rad(-60)), configs.cpp:43-44
  unsigned magic[SF_MAGIC_N];            // exact-division reciprocals used by the scan converter
  // hexagons (hexagon.cpp:13-48): vertex i and the edge normal (nx,ny) of edge i->i+1; [0]=big (200), [1]=small (40)
  double hex_px[2][6], hex_py[2][6], hex_nx[2][6], hex_ny[2][6];
  // resize tables (cv2 INTER_AREA 92x90 -> 84x84, rl/envs.py:29)
  int xt_cnt[84], xt_si[84][SF_MAX_TAPS];
  float xt_a[84][SF_MAX_TAPS];
  int yt_cnt[84], yt_si[84][SF_MAX_TAPS];
  float yt_a[84][SF_MAX_TAPS];
  SfTap xtap[84], ytap[84];                    // the same tables, packed for the device epilogue
  int col_out0[SF_NAT_W], col_out1[SF_NAT_W];  // first/last output column reading native column c
  int row_out0[SF_NAT_H], row_out1[SF_NAT_H];
  // explosion geometry relative to the fixed-point centre: per quad 4 corners (dx,dy)
  short exp_quad[SF_EXP_QUADS][8];
  unsigned char exp_colour[SF_EXP_STROKES];
  // frames
  alignas(16) unsigned char bg_nat[SF_NAT_H * SF_NAT_STRIDE];      // hexagons on black, native (read with 128-bit loads)
  alignas(16) unsigned char bg_obs[84 * 84];                       // resize(bg_nat)
  unsigned char fort_alpha[36][SF_FORT_W * SF_FORT_W]; // fortress wireframe coverage per sector angle
  unsigned char fexp_alpha[SF_EXP_LAYERS][SF_EXP_W * SF_EXP_W];  // fortress explosion, ordered layers
  unsigned char fexp_colour[SF_EXP_LAYERS][SF_EXP_W * SF_EXP_W];
  int fexp_layers;
  unsigned char text_alpha[10][SF_TEXT_H * SF_TEXT_W]; // digit d drawn in the slot owning each column
  unsigned char text_slot[SF_TEXT_W];                  // which of the 7 digit slots owns a strip column (255: none)
  unsigned char bar_alpha[SF_BAR_H];                   // per-row coverage of the vulnerability bar
  unsigned char colour_bar_bg, colour_bar_fg, colour_bar_kill, colour_text, colour_white;
  // wireframe models (wireframe.cpp:8-70): [0]=ship [1]=missile [2]=shell; line = from.x, from.y, to.x, to.y
  double wf_line[3][4][4];
  int wf_nlines[3];
  // sparse fortress sprites: lit pixels only (index into the 17x17 box, coverage), tight bounding rect
  unsigned short fort_list_idx[36][SF_FORT_LIST];
  unsigned short fort_list_xy[36][SF_FORT_LIST];  // the same pixels as native x | y<<8
  unsigned char fort_list_a[36][SF_FORT_LIST];
  int fort_list_n[36];
  unsigned char fort_rect[SF_FORT_STATES][4];  // x0,y0,x1,y1 native px of the lit pixels (state 36 = explosion)
  // Pre-resampled output for the static layers when no moving object overlaps them. bg_obs above already
  // contains score "0000000" and the empty vulnerability bar; these tables hold the 16-byte output chunks that
  // change with the fortress state / bar state (same base frame).
  int fort_chunk0, fort_nchunks;
  alignas(16) unsigned char obs_fort[SF_FORT_STATES][SF_FORT_CHUNKS * 16];
  alignas(16) unsigned char obs_bar[SF_BAR_STATES][(SF_OBS_CHUNKS - SF_BAR_CHUNK0) * 16];
  // the chunks of obs_fort[st] that differ from the default observation (index relative to fort_chunk0), 255-terminated
  unsigned char fort_sparse[SF_FORT_STATES][64];
  unsigned char fort_sparse_n[SF_FORT_STATES];
  SfExpPhase exp_phase[256];
  alignas(16) SfHot hot;
  int text_guard_row;  // moving rects with y0 <= this native row force the general text path
  int bar_guard_row;   // moving rects with y1 >= this native row force the general bar path
};

// builds the tables on the host; returns 0 on success, else a message in err
int sf_build_tables(SfTables* t, char* err, int errcap);
// the same with the score digits taken from glyph masks (sf_set_glyph_masks, include/sf_b200.h) instead of the built-in
// 7-segment face: alpha[10][SF_TEXT_H * SF_TEXT_W], slot[SF_TEXT_W]; both NULL == sf_build_tables
int sf_build_tables_glyphs(SfTables* t, char* err, int errcap, const unsigned char* alpha, const unsigned char* slot);
