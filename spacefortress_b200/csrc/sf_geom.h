// sf_geom.h — geometry shared by the device rasteriser and the host-side static-table builder.
//
// Frame model (the contract both sides implement; DESIGN.md §"Frame model"):
//  * path points go through the cairo CTM in double precision with cairo_matrix_multiply /
//    cairo_matrix_transform_point operation order and are stored as 24.8 fixed point, round-to-nearest-even
//    (reference call sites: draw.cpp:82-100, 256-261);
//  * a stroked segment is the quad  p0+o, p1+o, p1-o, p0-o  with o = fixed(half_width * unit normal);
//  * coverage: 15 sub-rows per pixel row, exact box filter in x at 1/256 px, alpha=(34*L+256)>>9;
//  * blend: d = mul8(c,a) + mul8(d,255-a).
// All double arithmetic is written with explicit round-to-nearest single operations (no FMA contraction)
// so that host (tables) and device (per-frame sprites) produce identical bits.
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define SF_HD __host__ __device__ __forceinline__
#else
#define SF_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define SF_DMUL(a, b) __dmul_rn((a), (b))
#define SF_DADD(a, b) __dadd_rn((a), (b))
#define SF_DSUB(a, b) __dsub_rn((a), (b))
#define SF_DDIV(a, b) __ddiv_rn((a), (b))
#define SF_DSQRT(a) __dsqrt_rn((a))
#define SF_RINT_I(a) __double2int_rn((a))
#else
// host: this translation unit is compiled with -ffp-contract=off
#define SF_DMUL(a, b) ((a) * (b))
#define SF_DADD(a, b) ((a) + (b))
#define SF_DSUB(a, b) ((a) - (b))
#define SF_DDIV(a, b) ((a) / (b))
#define SF_DSQRT(a) sqrt((a))
#define SF_RINT_I(a) ((int)nearbyint((a)))
#endif

#include "sf_tables.h"
#define SF_GRID_Y 15
#define SF_CTM_SCALE 0.2   // 90/450 == 92/460 == the double 0.2 (ssf_env.py:50,57-58)
#define SF_CTM_X0 (-26.0)  // fl(-130*0.2)
#define SF_CTM_Y0 (-16.0)  // fl(-80*0.2)
#define SF_FORT_X 355.0    // game.cpp:38-39
#define SF_FORT_Y 315.0

struct SfPt { int x, y; };           // 24.8 device coordinates
struct SfQuad { SfPt p[4]; };        // convex, consistent orientation

SF_HD int sf_to_fixed(double v) { return SF_RINT_I(SF_DMUL(v, 256.0)); }

// user point -> device fixed, base CTM only (hexagons, boxes, text, explosion centres)
SF_HD SfPt sf_xform_base(double ux, double uy) {
  SfPt p;
  p.x = sf_to_fixed(SF_DADD(SF_DMUL(SF_CTM_SCALE, ux), SF_CTM_X0));
  p.y = sf_to_fixed(SF_DADD(SF_DMUL(SF_CTM_SCALE, uy), SF_CTM_Y0));
  return p;
}

// CTM after cairo_translate(pos) and cairo_rotate(angle) (draw.cpp:85-86): c,s = cos/sin(deg2rad(int angle))
struct SfWireXf { double xx, yx, xy, yy, x0, y0; };

SF_HD SfWireXf sf_wire_xf(double px, double py, double c, double s) {
  SfWireXf m;
  m.x0 = SF_DADD(SF_DMUL(px, SF_CTM_SCALE), SF_CTM_X0);  // tx*xx + ty*xy(=0) + x0
  m.y0 = SF_DADD(SF_DMUL(py, SF_CTM_SCALE), SF_CTM_Y0);
  m.xx = SF_DMUL(c, SF_CTM_SCALE);
  m.yx = SF_DMUL(s, SF_CTM_SCALE);
  m.xy = SF_DMUL(-s, SF_CTM_SCALE);
  m.yy = SF_DMUL(c, SF_CTM_SCALE);
  return m;
}

SF_HD SfPt sf_xform_wire(const SfWireXf& m, double mx, double my) {
  SfPt p;
  p.x = sf_to_fixed(SF_DADD(SF_DADD(SF_DMUL(m.xx, mx), SF_DMUL(m.xy, my)), m.x0));
  p.y = sf_to_fixed(SF_DADD(SF_DADD(SF_DMUL(m.yx, mx), SF_DMUL(m.yy, my)), m.y0));
  return p;
}

// unit direction of a fixed-point segment (cairo normalize_slope) ; returns false if degenerate
SF_HD bool sf_unit_dir(SfPt a, SfPt b, double& ux, double& uy) {
  double dx = SF_DMUL((double)(b.x - a.x), 0.00390625), dy = SF_DMUL((double)(b.y - a.y), 0.00390625);  // exact: /256
  if (dx == 0.0 && dy == 0.0) return false;
  if (dx == 0.0) { ux = 0.0; uy = dy > 0 ? 1.0 : -1.0; }
  else if (dy == 0.0) { uy = 0.0; ux = dx > 0 ? 1.0 : -1.0; }
  else {
    double mag = SF_DSQRT(SF_DADD(SF_DMUL(dx, dx), SF_DMUL(dy, dy)));
    ux = SF_DDIV(dx, mag); uy = SF_DDIV(dy, mag);
  }
  return true;
}

// half line width in device px: (3/2 user units) * 0.2   (ssf_env.py:50 ls=3; draw.cpp:261)
#define SF_HALF_WIDTH_DEV (1.5 * SF_CTM_SCALE)

SF_HD bool sf_stroke_quad(SfPt a, SfPt b, SfQuad& q) {
  double ux, uy;
  if (!sf_unit_dir(a, b, ux, uy)) return false;
  const double hw = SF_DMUL(1.5, SF_CTM_SCALE);
  int ox = sf_to_fixed(SF_DMUL(-uy, hw)), oy = sf_to_fixed(SF_DMUL(ux, hw));
  q.p[0].x = a.x + ox; q.p[0].y = a.y + oy;
  q.p[1].x = b.x + ox; q.p[1].y = b.y + oy;
  q.p[2].x = b.x - ox; q.p[2].y = b.y - oy;
  q.p[3].x = a.x - ox; q.p[3].y = a.y - oy;
  return true;
}

SF_HD int sf_grid_y(int yfixed) { return (yfixed * SF_GRID_Y + 128) >> 8; }  // round-half-up to 1/15 px
SF_HD unsigned sf_mul8(unsigned a, unsigned b) { unsigned t = a * b + 0x80u; return (t + (t >> 8)) >> 8; }
SF_HD unsigned sf_blend(unsigned d, unsigned colour, unsigned a) {
  if (a >= 255u) return colour;
  unsigned v = sf_mul8(colour, a) + sf_mul8(d, 255u - a);
  return v > 255u ? 255u : v;
}
SF_HD unsigned sf_len_to_alpha(unsigned len) { return (34u * len + 256u) >> 9; }  // len in 1/256 px summed over 15 sub-rows

// wireframe models (wireframe.cpp:8-70): lines as from.x, from.y, to.x, to.y
#define SF_WF_SHIP_LINES 3
#define SF_WF_FORTRESS_LINES 4
#define SF_WF_MISSILE_LINES 3
#define SF_WF_SHELL_LINES 4
