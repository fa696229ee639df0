// sf_render.cuh — warp-cooperative rasteriser: one warp draws one env's frame into shared memory
// (native 92x90 tile), resamples the touched region to 84x84 with cv2's INTER_AREA arithmetic and
// streams the observation out with 128-bit stores. Replaces drawGameStateScaled (draw.cpp:256-270),
// the RGBA2GRAY conversion (ssf_env.py:205, identity on grey input) and cv2.resize (rl/envs.py:29).
//
// Draw order and semantics follow draw.cpp:227-269:
//   black, 2 hexagons [static: bg_nat]  ->  ship wireframe | ship explosion  ->  fortress wireframe |
//   fortress explosion [static sprites]  ->  missiles  ->  shells further than 21 from the fortress
//   ->  score digits [static glyph strip]  ->  vulnerability bar.
// Moving strokes (ship, missiles, shells, ship explosion) are scan-converted on the fly with the
// model of sf_geom.h: per stroke, lanes own (pixel-row, sub-row) samples, evaluate every quad's span
// with exact integer edge stepping, merge overlapping spans (non-zero winding == union for equally
// oriented convex quads) and accumulate 1/256-px span lengths into a shared-memory cell array.
#pragma once
#include "sf_geom.h"
#include "sf_state.cuh"
#include "sf_tables.h"

#define SF_ACC_CELLS 1024
#define SF_MAX_STROKE_QUADS 16
#define SF_MAX_RECTS 40

// per-warp shared memory
struct __align__(16) SfWarpSmem {
  unsigned char nat[SF_NAT_H * SF_NAT_STRIDE];  // 8832 B
  unsigned char out[84 * 84];                   // 7056 B
  int acc[SF_ACC_CELLS];                        // span-length accumulators of the current stroke
  int4 edge[SF_MAX_STROKE_QUADS * 4];           // x_top, y_top(grid), dy(grid), dx
  unsigned edge_m[SF_MAX_STROKE_QUADS * 4];     // magic reciprocal of dy
  SfQuad quad[SF_MAX_STROKE_QUADS];
  int4 rect[SF_MAX_RECTS];                      // dirty rectangles, native px inclusive
  int nrect;
  int pad[3];
};

__constant__ double c_wf_ship[3][4] = {{-18, 0, 18, 0}, {-18, 18, 0, 0}, {0, 0, -18, -18}};       // wireframe.cpp:39-53
__constant__ double c_wf_missile[3][4] = {{0, 0, -25, 0}, {0, 0, -5, 5}, {0, 0, -5, -5}};          // wireframe.cpp:11-22
__constant__ double c_wf_shell[4][4] = {{-8, 0, 0, -6}, {0, -6, 16, 0}, {16, 0, 0, 6}, {0, 6, -8, 0}};  // wireframe.cpp:24-37

__device__ __forceinline__ int sf_warp_min(int v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ int sf_warp_max(int v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ void sf_add_rect(SfWarpSmem& W, int lane, int x0, int y0, int x1, int y1) {
  if (lane == 0) {
    int n = W.nrect;
    if (n < SF_MAX_RECTS) { W.rect[n] = make_int4(x0, y0, x1, y1); W.nrect = n + 1; }
    else {  // overflow: grow the last rectangle to the union (still correct, just more resampling)
      int4 r = W.rect[SF_MAX_RECTS - 1];
      W.rect[SF_MAX_RECTS - 1] = make_int4(min(r.x, x0), min(r.y, y0), max(r.z, x1), max(r.w, y1));
    }
  }
}

// exact floor(m*dx/dy) for 0 <= m < dy
__device__ __forceinline__ int sf_edge_x(int4 E, unsigned M, int m) {
  int dx = E.w, dy = E.z;
  unsigned adx = (unsigned)abs(dx);
  unsigned n = (unsigned)m * adx + (dx < 0 ? (unsigned)(dy - 1) : 0u);
  unsigned q = M ? __umulhi(n, M) : (dy > 1 ? n / (unsigned)dy : 0u);
  return E.x + (dx < 0 ? -(int)q : (int)q);
}

// Scan-convert the stroke made of quads W.quad[0..nq) (quads in groups of 4 may overlap inside a group;
// different groups must be disjoint) and blend `colour` into the native tile.
__device__ __noinline__ void sf_raster_stroke(SfWarpSmem& W, int lane, int nq, unsigned colour) {
  // ---- edges ----
  int smin = 1 << 30, smax = -(1 << 30), xmin = 1 << 30, xmax = -(1 << 30);
  for (int e = lane; e < nq * 4; e += 32) {
    SfPt a = W.quad[e >> 2].p[e & 3], b = W.quad[e >> 2].p[(e + 1) & 3];
    int ga = sf_grid_y(a.y), gb = sf_grid_y(b.y);
    xmin = min(xmin, min(a.x, b.x)); xmax = max(xmax, max(a.x, b.x));
    int4 E; unsigned M = 0;
    if (ga == gb) { E = make_int4(0, 0, 0, 0); }
    else {
      if (ga < gb) E = make_int4(a.x, ga, gb - ga, b.x - a.x); else E = make_int4(b.x, gb, ga - gb, a.x - b.x);
      smin = min(smin, E.y); smax = max(smax, E.y + E.z);
      unsigned dy = (unsigned)E.z, adx = (unsigned)abs(E.w);
      // magic multiply is exact while (dy*adx + dy) * dy < 2^32; otherwise fall back to a real division
      if (dy > 1 && (unsigned long long)(dy * (unsigned long long)adx + dy) * dy < (1ull << 32)) M = (unsigned)((1ull << 32) / dy) + 1u;
    }
    W.edge[e] = E; W.edge_m[e] = M;
  }
  smin = sf_warp_min(smin); smax = sf_warp_max(smax); xmin = sf_warp_min(xmin); xmax = sf_warp_max(xmax);
  if (smin >= smax) return;
  // floor division by 15 of possibly negative grid rows
  int py0 = (smin >= 0) ? smin / SF_GRID_Y : -((-smin + SF_GRID_Y - 1) / SF_GRID_Y);
  int py1 = (smax - 1 >= 0) ? (smax - 1) / SF_GRID_Y : -((-(smax - 1) + SF_GRID_Y - 1) / SF_GRID_Y);
  py0 = max(py0, 0); py1 = min(py1, SF_NAT_H - 1);
  int cx0 = max(xmin >> 8, 0), cx1 = min((xmax - 1) >> 8, SF_NAT_W - 1);
  if (py0 > py1 || cx0 > cx1) return;  // entirely off the surface
  const int w = cx1 - cx0 + 1, h = py1 - py0 + 1;
  const int xlo = cx0 << 8, xhi = (cx1 + 1) << 8;
  const int hc = min(h, SF_ACC_CELLS / w);
  __syncwarp();
  for (int r0 = 0; r0 < h; r0 += hc) {
    const int hh = min(hc, h - r0);
    const float inv_hh = 1.0f / (float)hh;
    const int items = hh * SF_GRID_Y;
    for (int it0 = 0; it0 < items; it0 += 32) {
      int it = it0 + lane;
      if (it < items) {
        int k = __float2int_rz(((float)it + 0.5f) * inv_hh);  // it / hh (exact: |frac| >= 0.5/hh)
        int r = it - k * hh;
        int s = (py0 + r0 + r) * SF_GRID_Y + k;
        for (int g = 0; g < nq; g += 4) {
          unsigned key[4];
#pragma unroll
          for (int j = 0; j < 4; j++) {
            key[j] = 0xFFFFFFFFu;
            if (g + j < nq) {
              int lo = 1 << 30, hi = -(1 << 30);
#pragma unroll
              for (int c = 0; c < 4; c++) {
                int4 E = W.edge[(g + j) * 4 + c];
                int m = s - E.y;
                if ((unsigned)m < (unsigned)E.z) {
                  int x = sf_edge_x(E, W.edge_m[(g + j) * 4 + c], m);
                  lo = min(lo, x); hi = max(hi, x);
                }
              }
              lo = max(lo, xlo); hi = min(hi, xhi);
              if (lo < hi) key[j] = ((unsigned)(lo - xlo) << 16) | (unsigned)(hi - xlo);
            }
          }
          // sort the 4 spans by start (5-comparator network), then emit each span minus the union of its predecessors
#define SF_CE(a, b) { unsigned lo_ = min(key[a], key[b]), hi_ = max(key[a], key[b]); key[a] = lo_; key[b] = hi_; }
          SF_CE(0, 1) SF_CE(2, 3) SF_CE(0, 2) SF_CE(1, 3) SF_CE(1, 2)
#undef SF_CE
          int reach = 0;
#pragma unroll
          for (int j = 0; j < 4; j++) {
            if (key[j] == 0xFFFFFFFFu) break;
            int a = max((int)(key[j] >> 16), reach), b = (int)(key[j] & 0xFFFFu);
            reach = max(reach, b);
            int cell = a >> 8;
            while (a < b) {
              int e = min(b, (cell + 1) << 8);
              atomicAdd(&W.acc[r * w + cell], e - a);
              a = e; cell++;
            }
          }
        }
      }
    }
    __syncwarp();
    const float inv_w = 1.0f / (float)w;
    for (int idx = lane; idx < hh * w; idx += 32) {
      int L = W.acc[idx];
      if (L) {
        W.acc[idx] = 0;
        int r = __float2int_rz(((float)idx + 0.5f) * inv_w), c = idx - r * w;
        unsigned char* px = &W.nat[(py0 + r0 + r) * SF_NAT_STRIDE + cx0 + c];
        *px = (unsigned char)sf_blend(*px, colour, sf_len_to_alpha((unsigned)L));
      }
    }
    __syncwarp();
  }
  sf_add_rect(W, lane, cx0, py0, cx1, py1);
}

// R3 drawWireFrame (draw.cpp:82-100): lanes < nlines build one stroked quad each
__device__ __forceinline__ void sf_wireframe(SfWarpSmem& W, int lane, const SfTables* T, const double (*lines)[4], int nlines,
                                             double px, double py, int angle, unsigned colour) {
  // quick cull: every model fits in a 37-unit radius (7.4 px) around its origin
  double dxv = SF_DADD(SF_DMUL(px, SF_CTM_SCALE), SF_CTM_X0), dyv = SF_DADD(SF_DMUL(py, SF_CTM_SCALE), SF_CTM_Y0);
  if (dxv < -9.0 || dxv > SF_NAT_W + 9.0 || dyv < -9.0 || dyv > SF_NAT_H + 9.0) return;
  if (lane < nlines) {
    SfWireXf m = sf_wire_xf(px, py, T->cos_deg[angle], T->sin_deg[angle]);
    SfQuad q;
    SfPt a = sf_xform_wire(m, lines[lane][0], lines[lane][1]), b = sf_xform_wire(m, lines[lane][2], lines[lane][3]);
    if (!sf_stroke_quad(a, b, q)) { q.p[0] = q.p[1] = q.p[2] = q.p[3] = a; }  // degenerate: no area
    W.quad[lane] = q;
  }
  __syncwarp();
  sf_raster_stroke(W, lane, nlines, colour);
}

// R5 drawExplosion (draw.cpp:116-145): 84 arcs, each its own stroke, then the r=7 circle
__device__ __noinline__ void sf_explosion(SfWarpSmem& W, int lane, const SfTables* T, double px, double py) {
  SfPt c = sf_xform_base(px, py);
  for (int s = 0; s < SF_EXP_STROKES; s++) {
    int nq = (s == SF_EXP_STROKES - 1) ? 16 : 1;
    if (lane < nq) {
      const short* o = T->exp_quad[s + lane];  // the circle's 16 quads start at index 84 == s
      SfQuad q;
#pragma unroll
      for (int j = 0; j < 4; j++) { q.p[j].x = c.x + o[2 * j]; q.p[j].y = c.y + o[2 * j + 1]; }
      W.quad[lane] = q;
    }
    __syncwarp();
    sf_raster_stroke(W, lane, nq, T->exp_colour[s]);
  }
}

// one output pixel of cv2 INTER_AREA (float accumulation in table order, round-half-even)
__device__ __forceinline__ unsigned char sf_resample(const SfWarpSmem& W, const SfTables* T, int i, int j) {
  float sum = 0.f;
  int ny = T->yt_cnt[i], nx = T->xt_cnt[j];
  for (int ky = 0; ky < ny; ky++) {
    const unsigned char* S = &W.nat[T->yt_si[i][ky] * SF_NAT_STRIDE];
    float buf = 0.f;
    for (int kx = 0; kx < nx; kx++) buf = __fadd_rn(buf, __fmul_rn((float)S[T->xt_si[j][kx]], T->xt_a[j][kx]));
    float v = __fmul_rn(T->yt_a[i][ky], buf);
    sum = ky == 0 ? v : __fadd_rn(sum, v);
  }
  return (unsigned char)__float2int_rn(sum);
}

// once per warp at kernel start: the span accumulators must start at zero (every stroke leaves them zeroed)
__device__ __forceinline__ void sf_warp_smem_init(SfWarpSmem& W, int lane) {
  for (int k = lane; k < SF_ACC_CELLS; k += 32) W.acc[k] = 0;
  if (lane == 0) W.nrect = 0;
  __syncwarp();
}

struct SfRenderIn {  // warp-uniform view of one env
  int env;
  unsigned core, pmask;
  double px, py;
  int points_i, vuln;
  bool kill_bar;  // vuln > 10 && vulnerability timer < 250 (draw.cpp:268)
};

// Draw env `in` and write its observation. obs84: 84*84 bytes (or NULL), nat_out: 92*90 bytes (or NULL).
__device__ __noinline__ void sf_render_env(const SfDev& D, SfWarpSmem& W, int lane, const SfRenderIn& in,
                                           unsigned char* __restrict__ obs84, unsigned char* __restrict__ nat_out) {
  const SfTables* T = D.tab;
  const int np = D.n_pad;
  // ---- background ----
  {
    const int4* src = reinterpret_cast<const int4*>(T->bg_nat);
    int4* dst = reinterpret_cast<int4*>(W.nat);
    for (int k = lane; k < SF_NAT_H * SF_NAT_STRIDE / 16; k += 32) dst[k] = __ldg(&src[k]);
    if (lane == 0) W.nrect = 0;
  }
  __syncwarp();
  // ---- ship (draw.cpp:233-237) ----
  if (in.core & SF_CORE_SHIP_ALIVE) sf_wireframe(W, lane, T, c_wf_ship, 3, in.px, in.py, (int)(in.core & SF_CORE_ANGLE_MASK), T->colour_white);
  else sf_explosion(W, lane, T, in.px, in.py);
  // ---- fortress (draw.cpp:238-242): static sprites at (355,315) ----
  if (in.core & SF_CORE_FORT_ALIVE) {
    const unsigned char* A = T->fort_alpha[(in.core >> SF_CORE_FANG_SHIFT) & 63u];
    for (int idx = lane; idx < SF_FORT_W * SF_FORT_W; idx += 32) {
      unsigned a = A[idx];
      if (a) {
        int r = idx / SF_FORT_W, c = idx - r * SF_FORT_W;
        unsigned char* px = &W.nat[(SF_FORT_Y0 + r) * SF_NAT_STRIDE + SF_FORT_X0 + c];
        *px = (unsigned char)sf_blend(*px, T->colour_white, a);
      }
    }
    sf_add_rect(W, lane, SF_FORT_X0, SF_FORT_Y0, SF_FORT_X0 + SF_FORT_W - 1, SF_FORT_Y0 + SF_FORT_W - 1);
  } else {
    for (int idx = lane; idx < SF_EXP_W * SF_EXP_W; idx += 32) {
      int r = idx / SF_EXP_W, c = idx - r * SF_EXP_W;
      unsigned char* px = &W.nat[(SF_FEXP_Y0 + r) * SF_NAT_STRIDE + SF_FEXP_X0 + c];
      unsigned v = *px;
      for (int l = 0; l < T->fexp_layers; l++) {
        unsigned a = T->fexp_alpha[l][idx];
        if (!a) break;
        v = sf_blend(v, T->fexp_colour[l][idx], a);
      }
      *px = (unsigned char)v;
    }
    sf_add_rect(W, lane, SF_FEXP_X0, SF_FEXP_Y0, SF_FEXP_X0 + SF_EXP_W - 1, SF_FEXP_Y0 + SF_EXP_W - 1);
  }
  __syncwarp();
  // ---- missiles (draw.cpp:243-247), slot order ----
  for (unsigned m = in.pmask & SF_PMASK_MISSILES; m; m &= m - 1) {
    int s = __ffs(m) - 1;
    double2 p = D.mpos[(size_t)s * np + in.env];
    int ang = D.mang[(size_t)s * np + in.env];
    sf_wireframe(W, lane, T, c_wf_missile, 3, p.x, p.y, ang, T->colour_white);
  }
  // ---- shells (draw.cpp:248-253): hidden within 21 units of the fortress (quirk Q9); int angle (Q10) ----
  for (unsigned m = (in.pmask >> SF_PMASK_SHELL_SHIFT) & 0xFu; m; m &= m - 1) {
    int s = __ffs(m) - 1;
    double2 p = D.spos[(size_t)s * np + in.env];
    double dx = SF_DSUB(p.x, SF_FORT_X), dy = SF_DSUB(p.y, SF_FORT_Y);
    if (SF_DSQRT(SF_DADD(SF_DMUL(dx, dx), SF_DMUL(dy, dy))) > 21.0) {
      int ang = __double2int_rz(D.sang[(size_t)s * np + in.env]);
      if (ang >= 360) ang -= 360;
      sf_wireframe(W, lane, T, c_wf_shell, 4, p.x, p.y, ang, T->colour_white);
    }
  }
  // ---- score digits (draw.cpp:160-173,267): "%07d" of (int)mPoints ----
  {
    int pts = min(max(in.points_i, 0), 9999999);
    for (int idx = lane; idx < SF_TEXT_H * SF_TEXT_W; idx += 32) {
      int r = idx / SF_TEXT_W, c = idx - r * SF_TEXT_W;
      int slot = T->text_slot[c];
      if (slot < 7) {
        int div = 1;
        for (int k = slot; k < 6; k++) div *= 10;
        unsigned a = T->text_alpha[(pts / div) % 10][idx];
        if (a) {
          unsigned char* px = &W.nat[(SF_TEXT_Y0 + r) * SF_NAT_STRIDE + SF_TEXT_X0 + c];
          *px = (unsigned char)sf_blend(*px, T->colour_text, a);
        }
      }
    }
    sf_add_rect(W, lane, SF_TEXT_X0, SF_TEXT_Y0, SF_TEXT_X0 + SF_TEXT_W - 1, SF_TEXT_Y0 + SF_TEXT_H - 1);
  }
  // ---- vulnerability bar (draw.cpp:207-225,268) ----
  {
    int filled = 4 * min(in.vuln, 10);  // 20 user units per step = 4 px
    unsigned fg = in.kill_bar ? T->colour_bar_kill : T->colour_bar_fg;
    for (int idx = lane; idx < SF_BAR_H * SF_BAR_W; idx += 32) {
      int r = idx / SF_BAR_W, c = idx - r * SF_BAR_W;
      unsigned a = T->bar_alpha[r];
      unsigned char* px = &W.nat[(SF_BAR_Y0 + r) * SF_NAT_STRIDE + SF_BAR_X0 + c];
      unsigned v = sf_blend(*px, T->colour_bar_bg, a);
      if (c < filled) v = sf_blend(v, fg, a);
      *px = (unsigned char)v;
    }
    sf_add_rect(W, lane, SF_BAR_X0, SF_BAR_Y0, SF_BAR_X0 + SF_BAR_W - 1, SF_BAR_Y0 + SF_BAR_H - 1);
  }
  __syncwarp();
  // ---- native output (SSF_Env.step returns the 92x90 frame) ----
  if (nat_out) {
    for (int idx = lane; idx < SF_NAT_H * SF_NAT_W / 2; idx += 32) {
      int r = idx / (SF_NAT_W / 2), c = (idx - r * (SF_NAT_W / 2)) * 2;
      *reinterpret_cast<uchar2*>(&nat_out[r * SF_NAT_W + c]) = *reinterpret_cast<const uchar2*>(&W.nat[r * SF_NAT_STRIDE + c]);
    }
  }
  // ---- 84x84 observation: static background + resampled dirty rectangles ----
  if (obs84) {
    const int4* src = reinterpret_cast<const int4*>(T->bg_obs);
    int4* dst = reinterpret_cast<int4*>(W.out);
    for (int k = lane; k < 84 * 84 / 16; k += 32) dst[k] = __ldg(&src[k]);
    __syncwarp();
    const int nr = W.nrect;
    for (int q = 0; q < nr; q++) {
      int4 R = W.rect[q];
      int j0 = T->col_out0[R.x], j1 = T->col_out1[R.z], i0 = T->row_out0[R.y], i1 = T->row_out1[R.w];
      int ow = j1 - j0 + 1, cnt = ow * (i1 - i0 + 1);
      float inv_ow = 1.0f / (float)ow;
      for (int idx = lane; idx < cnt; idx += 32) {
        int r = __float2int_rz(((float)idx + 0.5f) * inv_ow), c = idx - r * ow;
        W.out[(i0 + r) * 84 + j0 + c] = sf_resample(W, T, i0 + r, j0 + c);
      }
    }
    __syncwarp();
    int4* g = reinterpret_cast<int4*>(obs84);
    const int4* o = reinterpret_cast<const int4*>(W.out);
    for (int k = lane; k < 84 * 84 / 16; k += 32) __stcs(&g[k], o[k]);  // streaming store: obs is write-once
  }
  __syncwarp();
}
