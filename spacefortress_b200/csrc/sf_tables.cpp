// sf_tables.cpp — host-side builder of the static render/physics tables (see sf_tables.h).
// Compiled with -ffp-contract=off. Reference call sites restated here:
//   hexagons        hexagon.cpp:13-35, draw.cpp:102-114,230-231  (closed path, miter joins)
//   fortress        wireframe.cpp:55-67, draw.cpp:238-242, game.cpp:38-40,206 (36 sector angles)
//   explosion       draw.cpp:116-145
//   score digits    draw.cpp:147-173 (font dependent; 7-segment face, DESIGN.md "Frame model" M7)
//   vulnerability   draw.cpp:207-225
//   resize          rl/envs.py:28-30 (cv2 INTER_AREA; OpenCV ResizeArea_ tables)
#include "sf_tables.h"

#include <math.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "sf_geom.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace {

struct HEdge { int x1, y1, x2, y2, dir; };  // fixed-point, y1 < y2

struct HPoly {
  std::vector<HEdge> e;
  void edge(SfPt a, SfPt b) {
    if (a.y == b.y) return;
    HEdge h;
    if (a.y < b.y) h = {a.x, a.y, b.x, b.y, 1}; else h = {b.x, b.y, a.x, a.y, -1};
    e.push_back(h);
  }
  void contour(const SfPt* p, int n) { for (int i = 0; i < n; i++) edge(p[i], p[(i + 1) % n]); }
  void quad(const SfQuad& q) { contour(q.p, 4); }
};

long long fdiv(long long a, long long b) {  // floor division, b > 0
  long long q = a / b;
  if ((a % b) < 0) q -= 1;
  return q;
}

// Sub-sampled non-zero-winding fill; returns per-pixel summed span length (1/256 px over 15 sub-rows).
void coverage(const HPoly& poly, std::vector<int>& len) {
  len.assign(SF_NAT_W * SF_NAT_H, 0);
  struct Live { int ytop, ybot; long long x1, dx, dy; int dir; };
  std::vector<Live> live;
  for (const HEdge& h : poly.e) {
    int g1 = sf_grid_y(h.y1), g2 = sf_grid_y(h.y2);
    if (g1 >= g2) continue;
    live.push_back({g1, g2, h.x1, (long long)h.x2 - h.x1, (long long)g2 - g1, h.dir});
  }
  std::vector<std::pair<int, int>> cross;
  for (int s = 0; s < SF_NAT_H * SF_GRID_Y; s++) {
    cross.clear();
    for (const Live& l : live)
      if (l.ytop <= s && s < l.ybot) cross.push_back({(int)(l.x1 + fdiv((long long)(s - l.ytop) * l.dx, l.dy)), l.dir});
    if (cross.size() < 2) continue;
    std::stable_sort(cross.begin(), cross.end(), [](const std::pair<int, int>& a, const std::pair<int, int>& b) { return a.first < b.first; });
    int wind = 0, start = 0;
    int* row = &len[(s / SF_GRID_Y) * SF_NAT_W];
    for (const auto& c : cross) {
      int prev = wind;
      wind += c.second;
      if (prev == 0 && wind != 0) start = c.first;
      else if (prev != 0 && wind == 0) {
        int a = std::max(start, 0), b = std::min(c.first, SF_NAT_W * 256);
        for (int px = a >> 8; px < SF_NAT_W && (px << 8) < b; px++) {
          int lo = std::max(a, px << 8), hi = std::min(b, (px << 8) + 256);
          if (hi > lo) row[px] += hi - lo;
        }
      }
    }
  }
}

void alpha_of(const HPoly& poly, std::vector<unsigned char>& alpha) {
  std::vector<int> len;
  coverage(poly, len);
  alpha.resize(len.size());
  for (size_t i = 0; i < len.size(); i++) alpha[i] = (unsigned char)sf_len_to_alpha((unsigned)len[i]);
}

void composite(unsigned char* img, int stride, const std::vector<unsigned char>& alpha, unsigned colour) {
  for (int y = 0; y < SF_NAT_H; y++)
    for (int x = 0; x < SF_NAT_W; x++) {
      unsigned a = alpha[y * SF_NAT_W + x];
      if (a) img[y * stride + x] = (unsigned char)sf_blend(img[y * stride + x], colour, a);
    }
}

unsigned colour8(double v) { return ((unsigned)(v * 65535.0 + 0.5)) >> 8; }

// closed polygon stroke with miter joins (cairo-path-stroke-polygon.c outer_join / inner_join)
void stroke_closed(HPoly& poly, const SfPt* v, int n) {
  std::vector<double> ux(n), uy(n);
  std::vector<SfPt> off(n);
  const double hw = SF_DMUL(1.5, SF_CTM_SCALE);
  for (int i = 0; i < n; i++) {
    sf_unit_dir(v[i], v[(i + 1) % n], ux[i], uy[i]);
    off[i].x = sf_to_fixed(SF_DMUL(-uy[i], hw));
    off[i].y = sf_to_fixed(SF_DMUL(ux[i], hw));
  }
  std::vector<SfPt> ccw, cw;
  for (int i = 0; i < n; i++) {
    int in = (i + n - 1) % n, out = i;
    SfPt P = v[i];
    SfPt in_ccw = {P.x + off[in].x, P.y + off[in].y}, in_cw = {P.x - off[in].x, P.y - off[in].y};
    SfPt out_ccw = {P.x + off[out].x, P.y + off[out].y}, out_cw = {P.x - off[out].x, P.y - off[out].y};
    bool inner_ccw = (ux[in] * uy[out] - uy[in] * ux[out]) > 0;
    SfPt ip = inner_ccw ? in_cw : in_ccw, op = inner_ccw ? out_cw : out_ccw;
    double x1 = ip.x / 256.0, y1 = ip.y / 256.0, x2 = op.x / 256.0, y2 = op.y / 256.0;
    double dx1 = ux[in], dy1 = uy[in], dx2 = ux[out], dy2 = uy[out];
    double my = ((x2 - x1) * dy1 * dy2 - y2 * dx2 * dy1 + y1 * dx1 * dy2) / (dx1 * dy2 - dx2 * dy1);
    double mx = fabs(dy1) >= fabs(dy2) ? (my - y1) * dx1 / dy1 + x1 : (my - y2) * dx2 / dy2 + x2;
    SfPt tip = {sf_to_fixed(mx), sf_to_fixed(my)};
    std::vector<SfPt>& outer = inner_ccw ? cw : ccw;
    std::vector<SfPt>& inner = inner_ccw ? ccw : cw;
    outer.push_back(inner_ccw ? in_cw : in_ccw); outer.push_back(tip); outer.push_back(inner_ccw ? out_cw : out_ccw);
    inner.push_back(inner_ccw ? in_ccw : in_cw); inner.push_back(P); inner.push_back(inner_ccw ? out_ccw : out_cw);
  }
  poly.contour(ccw.data(), (int)ccw.size());
  std::reverse(cw.begin(), cw.end());
  poly.contour(cw.data(), (int)cw.size());
}

void hexagon_points(int radius, SfPt* v) {  // hexagon.cpp:13-35
  double x1 = floor(355.0 - radius), x2 = floor(355.0 - radius * 0.5), x3 = floor(355.0 + radius * 0.5), x4 = floor(355.0 + radius);
  double y1 = 315, y2 = floor(315.0 - radius * sin(M_PI * 2 / 3)), y3 = floor(315.0 + radius * sin(M_PI * 2 / 3));
  double ux[6] = {x1, x2, x3, x4, x3, x2}, uy[6] = {y1, y2, y2, y1, y3, y3};
  for (int i = 0; i < 6; i++) v[i] = sf_xform_base(ux[i], uy[i]);
}

// cv2 INTER_AREA table for one axis (OpenCV computeResizeAreaTab)
struct Tap { int di, si; float a; };
std::vector<Tap> area_tab(int ssize, int dsize) {
  std::vector<Tap> t;
  double scale = (double)ssize / dsize;
  for (int d = 0; d < dsize; d++) {
    double f1 = d * scale, f2 = f1 + scale, cell = std::min(scale, ssize - f1);
    int s1 = (int)ceil(f1), s2 = (int)floor(f2);
    s2 = std::min(s2, ssize - 1);
    s1 = std::min(s1, s2);
    if (s1 - f1 > 1e-3) t.push_back({d, s1 - 1, (float)((s1 - f1) / cell)});
    for (int s = s1; s < s2; s++) t.push_back({d, s, (float)(1.0 / cell)});
    if (f2 - s2 > 1e-3) t.push_back({d, s2, (float)(std::min(std::min(f2 - s2, 1.0), cell) / cell)});
  }
  return t;
}

}  // namespace

static const double WF_FORTRESS[4][4] = {{0, 0, 36, 0}, {0, -18, 18, -18}, {18, -18, 18, 18}, {18, 18, 0, 18}};

int sf_build_tables(SfTables* t, char* err, int errcap) { return sf_build_tables_glyphs(t, err, errcap, nullptr, nullptr); }

int sf_build_tables_glyphs(SfTables* t, char* err, int errcap, const unsigned char* glyph_alpha, const unsigned char* glyph_slot) {
  memset(t, 0, sizeof(*t));
  for (int a = 0; a < 360; a++) {
    double r = (double)a * M_PI / 180;
    t->cos_deg[a] = cos(r);
    t->sin_deg[a] = sin(r);
  }
  {
    static const double ox[8] = {1, 1, 0, -1, -1, -1, 0, 1}, oy[8] = {0, 1, 1, 1, 0, -1, -1, -1};
    for (int k = 0; k < 8; k++) t->atan2_oct[k] = atan2(oy[k], ox[k]);
  }
  for (int h = 0; h < 2; h++) {  // hexagon.cpp:13-35 (vertices floored to integers), :40-41 (normals)
    int radius = h == 0 ? 200 : 40;
    double x1 = floor(355.0 - radius), x2 = floor(355.0 - radius * 0.5), x3 = floor(355.0 + radius * 0.5), x4 = floor(355.0 + radius);
    double y1 = 315, y2 = floor(315.0 - radius * sin(M_PI * 2 / 3)), y3 = floor(315.0 + radius * sin(M_PI * 2 / 3));
    double ux[6] = {x1, x2, x3, x4, x3, x2}, uy[6] = {y1, y2, y2, y1, y3, y3};
    for (int i = 0; i < 6; i++) {
      int j = (i + 1) % 6;
      t->hex_px[h][i] = ux[i]; t->hex_py[h][i] = uy[i];
      t->hex_nx[h][i] = -(uy[j] - uy[i]); t->hex_ny[h][i] = ux[j] - ux[i];
    }
  }
  t->magic[0] = 0; t->magic[1] = 0;  // dy == 1: the only sample has m == 0
  for (unsigned d = 2; d < SF_MAGIC_N; d++) t->magic[d] = 0xFFFFFFFFu / d + 1u;
  t->ship_start_vx = cos(-60.0 * M_PI / 180);
  t->ship_start_vy = sin(-60.0 * M_PI / 180);
  t->colour_bar_bg = (unsigned char)colour8(.33);
  t->colour_bar_fg = (unsigned char)colour8(.66);
  t->colour_bar_kill = (unsigned char)colour8(1.0);
  t->colour_text = (unsigned char)colour8(.5);
  t->colour_white = (unsigned char)colour8(1.0);

  // ---- resize tables ----
  {
    std::vector<Tap> xt = area_tab(SF_NAT_W, 84), yt = area_tab(SF_NAT_H, 84);
    for (int c = 0; c < SF_NAT_W; c++) { t->col_out0[c] = 1 << 20; t->col_out1[c] = -1; }
    for (int r = 0; r < SF_NAT_H; r++) { t->row_out0[r] = 1 << 20; t->row_out1[r] = -1; }
    for (const Tap& p : xt) {
      int& n = t->xt_cnt[p.di];
      if (n >= SF_MAX_TAPS) { snprintf(err, errcap, "resize x tab overflow"); return 1; }
      t->xt_si[p.di][n] = p.si; t->xt_a[p.di][n] = p.a; n++;
      t->col_out0[p.si] = std::min(t->col_out0[p.si], p.di); t->col_out1[p.si] = std::max(t->col_out1[p.si], p.di);
    }
    for (const Tap& p : yt) {
      int& n = t->yt_cnt[p.di];
      if (n >= SF_MAX_TAPS) { snprintf(err, errcap, "resize y tab overflow"); return 1; }
      t->yt_si[p.di][n] = p.si; t->yt_a[p.di][n] = p.a; n++;
      t->row_out0[p.si] = std::min(t->row_out0[p.si], p.di); t->row_out1[p.si] = std::max(t->row_out1[p.si], p.di);
    }
  }

  for (int k = 0; k < 84; k++) {
    t->xtap[k].si = t->xt_si[k][0]; t->xtap[k].cnt = t->xt_cnt[k];
    t->ytap[k].si = t->yt_si[k][0]; t->ytap[k].cnt = t->yt_cnt[k];
    for (int j = 0; j < SF_MAX_TAPS; j++) {
      t->xtap[k].a[j] = j < t->xt_cnt[k] ? t->xt_a[k][j] : 0.f;
      t->ytap[k].a[j] = j < t->yt_cnt[k] ? t->yt_a[k][j] : 0.f;
      if (j < t->xt_cnt[k] && t->xt_si[k][j] != t->xt_si[k][0] + j) { snprintf(err, errcap, "x taps not consecutive"); return 1; }
      if (t->xt_cnt[k] > 2) { snprintf(err, errcap, "x table has more than 2 taps"); return 1; }
      if (j < t->yt_cnt[k] && t->yt_si[k][j] != t->yt_si[k][0] + j) { snprintf(err, errcap, "y taps not consecutive"); return 1; }
    }
  }
  {
    static const double WF[3][4][4] = {
        {{-18, 0, 18, 0}, {-18, 18, 0, 0}, {0, 0, -18, -18}, {0, 0, 0, 0}},      // ship    wireframe.cpp:39-53
        {{0, 0, -25, 0}, {0, 0, -5, 5}, {0, 0, -5, -5}, {0, 0, 0, 0}},           // missile wireframe.cpp:11-22
        {{-8, 0, 0, -6}, {0, -6, 16, 0}, {16, 0, 0, 6}, {0, 6, -8, 0}}};         // shell   wireframe.cpp:24-37
    memcpy(t->wf_line, WF, sizeof(WF));
    t->wf_nlines[0] = 3; t->wf_nlines[1] = 3; t->wf_nlines[2] = 4;
  }

  std::vector<unsigned char> alpha;

  // ---- background: two hexagons on black (draw.cpp:262-263, 230-231) ----
  {
    for (int radius : {200, 40}) {
      SfPt v[6];
      hexagon_points(radius, v);
      HPoly poly;
      stroke_closed(poly, v, 6);
      alpha_of(poly, alpha);
      composite(t->bg_nat, SF_NAT_STRIDE, alpha, t->colour_white);
    }
  }

  // ---- fortress sprite per sector angle ----
  for (int k = 0; k < 36; k++) {
    int ang = 10 * k;
    SfWireXf m = sf_wire_xf(SF_FORT_X, SF_FORT_Y, t->cos_deg[ang], t->sin_deg[ang]);
    HPoly poly;
    for (int i = 0; i < SF_WF_FORTRESS_LINES; i++) {
      SfQuad q;
      if (sf_stroke_quad(sf_xform_wire(m, WF_FORTRESS[i][0], WF_FORTRESS[i][1]), sf_xform_wire(m, WF_FORTRESS[i][2], WF_FORTRESS[i][3]), q)) poly.quad(q);
    }
    alpha_of(poly, alpha);
    for (int y = 0; y < SF_NAT_H; y++)
      for (int x = 0; x < SF_NAT_W; x++) {
        unsigned a = alpha[y * SF_NAT_W + x];
        if (!a) continue;
        int sx = x - SF_FORT_X0, sy = y - SF_FORT_Y0;
        if (sx < 0 || sy < 0 || sx >= SF_FORT_W || sy >= SF_FORT_W) { snprintf(err, errcap, "fortress sprite box too small"); return 1; }
        t->fort_alpha[k][sy * SF_FORT_W + sx] = (unsigned char)a;
      }
  }

  // ---- explosion geometry (draw.cpp:116-145; DESIGN.md "Frame model" M6) ----
  {
    const double sc = SF_CTM_SCALE, hw = SF_DMUL(1.5, SF_CTM_SCALE);
    auto polar = [](double r, double deg, int& x, int& y) { x = sf_to_fixed(r * cos(deg * M_PI / 180)); y = sf_to_fixed(r * sin(deg * M_PI / 180)); };
    int q = 0, stroke = 0, ofs = 0;
    auto put = [&](double r_user, double a0, double a1) {
      int sx, sy, ex, ey, fsx, fsy, fex, fey;
      polar(r_user * sc, a0, sx, sy); polar(r_user * sc, a1, ex, ey);
      polar(hw, a0, fsx, fsy); polar(hw, a1, fex, fey);
      short* o = t->exp_quad[q++];
      o[0] = (short)(sx + fsx); o[1] = (short)(sy + fsy); o[2] = (short)(ex + fex); o[3] = (short)(ey + fey);
      o[4] = (short)(ex - fex); o[5] = (short)(ey - fey); o[6] = (short)(sx - fsx); o[7] = (short)(sy - fsy);
    };
    for (int radius = 15; radius < 70; radius += 8) {
      ofs += 3;
      for (int angle = 0; angle < 360; angle += 30) {
        put(radius, angle + ofs, angle + ofs + 10);
        t->exp_colour[stroke++] = (unsigned char)(radius < 60 ? colour8(.75) : colour8(.5));
      }
    }
    for (int k = 0; k < 16; k++) put(7, k * 22.5, (k + 1) * 22.5);
    t->exp_colour[stroke++] = (unsigned char)colour8(.75);
    if (q != SF_EXP_QUADS || stroke != SF_EXP_STROKES) { snprintf(err, errcap, "explosion table size"); return 1; }
  }

  // ---- the same quads scan-converted for each of the 256 y phases of the centre (SfExpPhase, sf_tables.h) ----
  {
    auto floordiv = [](int a, int b) { int q = a / b; return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q; };
    for (int phy = 0; phy < 256; phy++) {
      SfExpPhase& P = t->exp_phase[phy];
      memset(&P, 0, sizeof(P));
      int ni = 0, ns = 0;
      for (int q = 0; q < SF_EXP_QUADS; q++) {
        const short* o = t->exp_quad[q];
        int x[4], g[4];
        for (int k = 0; k < 4; k++) { x[k] = o[2 * k]; g[k] = sf_grid_y(phy + o[2 * k + 1]); }  // rows relative to 15 * (c.y >> 8)
        const int gmin = std::min(std::min(g[0], g[1]), std::min(g[2], g[3])), gmax = std::max(std::max(g[0], g[1]), std::max(g[2], g[3]));
        P.item0[q] = (unsigned short)ni;
        P.qxmin[q] = (short)std::min(std::min(x[0], x[1]), std::min(x[2], x[3]));
        P.row0[q] = (signed char)floordiv(gmin, SF_GRID_Y);
        int nd = 0, nu = 0;
        for (int k = 0; k < 4; k++) { const int d = g[(k + 1) & 3] - g[k]; nd += d > 0; nu += d < 0; }
        if (nd == 0 || nu == 0) continue;  // covers no sample row
        for (int r = floordiv(gmin, SF_GRID_Y); r <= floordiv(gmax - 1, SF_GRID_Y); r++) {
          if (ni >= SF_EXPT_ITEMS) { snprintf(err, errcap, "explosion phase table: items"); return 1; }
          SfExpItem& I = P.item[ni++];
          I.quad = (unsigned char)q; I.row = (signed char)r; I.span0 = (unsigned short)ns; I.n = 0;
          for (int sr = std::max(gmin, r * SF_GRID_Y); sr < std::min(gmax, (r + 1) * SF_GRID_Y); sr++) {
            // crossings of the edges that are live on sample row sr: x = xa + floor((sr - ga) * dx / dy)
            int lo = 1 << 30, hi = -(1 << 30);
            for (int k = 0; k < 4; k++) {
              int xa = x[k], ga = g[k], xb = x[(k + 1) & 3], gb = g[(k + 1) & 3];
              if (ga == gb) continue;
              if (ga > gb) { std::swap(xa, xb); std::swap(ga, gb); }
              if (sr < ga || sr >= gb) continue;
              const int xs = xa + (int)fdiv((long long)(sr - ga) * (xb - xa), gb - ga);
              lo = std::min(lo, xs); hi = std::max(hi, xs);
            }
            if (ns >= SF_EXPT_SPANS) { snprintf(err, errcap, "explosion phase table: spans"); return 1; }
            if (lo > hi) { lo = 0; hi = 0; }
            if (lo < hi && (lo < -13 * 256 || hi > 14 * 256 + 1 || r < -13 || r > 14)) { snprintf(err, errcap, "explosion phase table: box"); return 1; }
            if (((lo - P.qxmin[q]) >> 8) >= SF_EXPT_NC || (hi > lo && ((hi - 1 - P.qxmin[q]) >> 8) + 1 >= SF_EXPT_NC)) { snprintf(err, errcap, "explosion phase table: cells per item"); return 1; }
            P.span[ns][0] = (short)lo; P.span[ns][1] = (short)hi; ns++; I.n++;
          }
        }
      }
      P.n_items = ni; P.n_spans = ns;
    }
  }

  // ---- fortress explosion as ordered (alpha, colour) layers at the fixed fortress position ----
  {
    SfPt c = sf_xform_base(SF_FORT_X, SF_FORT_Y);
    std::vector<int> depth(SF_EXP_W * SF_EXP_W, 0);
    int q = 0;
    for (int s = 0; s < SF_EXP_STROKES; s++) {
      int nq = (s == SF_EXP_STROKES - 1) ? 16 : 1;
      HPoly poly;
      for (int k = 0; k < nq; k++, q++) {
        SfQuad qd;
        for (int j = 0; j < 4; j++) { qd.p[j].x = c.x + t->exp_quad[q][2 * j]; qd.p[j].y = c.y + t->exp_quad[q][2 * j + 1]; }
        poly.quad(qd);
      }
      alpha_of(poly, alpha);
      for (int y = 0; y < SF_NAT_H; y++)
        for (int x = 0; x < SF_NAT_W; x++) {
          unsigned a = alpha[y * SF_NAT_W + x];
          if (!a) continue;
          int sx = x - SF_FEXP_X0, sy = y - SF_FEXP_Y0;
          if (sx < 0 || sy < 0 || sx >= SF_EXP_W || sy >= SF_EXP_W) { snprintf(err, errcap, "explosion box too small"); return 1; }
          int& d = depth[sy * SF_EXP_W + sx];
          if (d >= SF_EXP_LAYERS) { snprintf(err, errcap, "explosion needs more than %d layers", SF_EXP_LAYERS); return 1; }
          t->fexp_alpha[d][sy * SF_EXP_W + sx] = (unsigned char)a;
          t->fexp_colour[d][sy * SF_EXP_W + sx] = t->exp_colour[s];
          d++;
          t->fexp_layers = std::max(t->fexp_layers, d);
        }
    }
  }

  // ---- score digits: masks rendered elsewhere by a real cairo + font, or the 7-segment face on the metrics of a 30-unit
  // monospace bold font ----
  if (glyph_alpha && glyph_slot) {
    memcpy(t->text_alpha, glyph_alpha, sizeof(t->text_alpha));
    memcpy(t->text_slot, glyph_slot, sizeof(t->text_slot));
    for (int c = 0; c < SF_TEXT_W; c++)
      if (t->text_slot[c] >= 7 && t->text_slot[c] != 255) { snprintf(err, errcap, "glyph slot of column %d is %d (0..6 or 255)", c, t->text_slot[c]); return 1; }
  } else {
    static const unsigned char SEG[10] = {0x3f, 0x06, 0x5b, 0x4f, 0x66, 0x6d, 0x7d, 0x07, 0x7f, 0x6f};
    static const double BOX[7][4] = {{3, 0, 12, 4}, {11, 0, 4, 13}, {11, 9, 4, 13}, {3, 18, 12, 4}, {3, 9, 4, 13}, {3, 0, 4, 13}, {3, 9, 12, 4}};
    const double x0 = 355 - 7 * 18 / 2.0, ytop = 97 - 22 / 2.0;
    memset(t->text_slot, 255, sizeof(t->text_slot));
    for (int slot = 0; slot < 7; slot++) {  // column ownership from the widest glyph (8)
      HPoly poly;
      for (int sg = 0; sg < 7; sg++) {
        double bx = x0 + 18 * slot + BOX[sg][0], by = ytop + BOX[sg][1];
        SfPt p[4] = {sf_xform_base(bx, by), sf_xform_base(bx + BOX[sg][2], by), sf_xform_base(bx + BOX[sg][2], by + BOX[sg][3]), sf_xform_base(bx, by + BOX[sg][3])};
        poly.contour(p, 4);
      }
      alpha_of(poly, alpha);
      for (int y = 0; y < SF_NAT_H; y++)
        for (int x = 0; x < SF_NAT_W; x++) if (alpha[y * SF_NAT_W + x]) {
          int sx = x - SF_TEXT_X0;
          if (sx < 0 || sx >= SF_TEXT_W) { snprintf(err, errcap, "text strip too small"); return 1; }
          if (t->text_slot[sx] != 255 && t->text_slot[sx] != slot) { snprintf(err, errcap, "text column shared by two digits"); return 1; }
          t->text_slot[sx] = (unsigned char)slot;
        }
    }
    for (int d = 0; d < 10; d++) {
      HPoly poly;  // the same digit in all 7 slots: every strip column belongs to exactly one slot
      for (int slot = 0; slot < 7; slot++)
        for (int sg = 0; sg < 7; sg++) if ((SEG[d] >> sg) & 1) {
          double bx = x0 + 18 * slot + BOX[sg][0], by = ytop + BOX[sg][1];
          SfPt p[4] = {sf_xform_base(bx, by), sf_xform_base(bx + BOX[sg][2], by), sf_xform_base(bx + BOX[sg][2], by + BOX[sg][3]), sf_xform_base(bx, by + BOX[sg][3])};
          poly.contour(p, 4);
        }
      alpha_of(poly, alpha);
      for (int y = 0; y < SF_NAT_H; y++)
        for (int x = 0; x < SF_NAT_W; x++) {
          unsigned a = alpha[y * SF_NAT_W + x];
          if (!a) continue;
          int sx = x - SF_TEXT_X0, sy = y - SF_TEXT_Y0;
          if (sx < 0 || sy < 0 || sx >= SF_TEXT_W || sy >= SF_TEXT_H) { snprintf(err, errcap, "text strip too small (%d,%d)", x, y); return 1; }
          t->text_alpha[d][sy * SF_TEXT_W + sx] = (unsigned char)a;
        }
    }
  }

  // ---- vulnerability bar rows: exact-area box coverage (draw.cpp:216-224) ----
  {
    SfPt a = sf_xform_base(355 - 100, 335 + 187), b = sf_xform_base(355 - 100 + 200, 335 + 187 + 10);
    if (a.x != SF_BAR_X0 * 256 || b.x != (SF_BAR_X0 + SF_BAR_W) * 256) { snprintf(err, errcap, "bar not pixel aligned in x"); return 1; }
    for (int r = 0; r < SF_BAR_H; r++) {
      int py = SF_BAR_Y0 + r;
      int lo = std::max(a.y, py * 256), hi = std::min(b.y, py * 256 + 256);
      unsigned area = hi > lo ? (unsigned)(hi - lo) * 256u : 0u;
      t->bar_alpha[r] = (unsigned char)((area * 255u + 32768u) >> 16);
    }
    if ((a.y >> 8) != SF_BAR_Y0 || ((b.y - 1) >> 8) != SF_BAR_Y0 + SF_BAR_H - 1) { snprintf(err, errcap, "bar rows"); return 1; }
  }

  // ---- output-space tables: INTER_AREA with the same float operation order as the device epilogue ----
  auto resample = [&](const std::vector<unsigned char>& nat, unsigned char* obs) {
    for (int i = 0; i < 84; i++)
      for (int j = 0; j < 84; j++) {
        float sum = 0.f;
        for (int ky = 0; ky < t->yt_cnt[i]; ky++) {
          float buf = 0.f;
          const unsigned char* S = &nat[t->yt_si[i][ky] * SF_NAT_STRIDE];
          for (int kx = 0; kx < t->xt_cnt[j]; kx++) buf += (float)S[t->xt_si[j][kx]] * t->xt_a[j][kx];
          sum = (ky == 0) ? t->yt_a[i][ky] * buf : sum + t->yt_a[i][ky] * buf;
        }
        obs[i * 84 + j] = (unsigned char)lrintf(sum);
      }
  };
  auto blend_px = [&](std::vector<unsigned char>& nat, int x, int y, unsigned colour, unsigned a) {
    unsigned char& d = nat[y * SF_NAT_STRIDE + x];
    d = (unsigned char)sf_blend(d, colour, a);
  };
  auto draw_text0 = [&](std::vector<unsigned char>& nat) {
    for (int i = 0; i < SF_TEXT_H * SF_TEXT_W; i++)
      if (t->text_slot[i % SF_TEXT_W] < 7 && t->text_alpha[0][i]) blend_px(nat, SF_TEXT_X0 + i % SF_TEXT_W, SF_TEXT_Y0 + i / SF_TEXT_W, t->colour_text, t->text_alpha[0][i]);
  };
  auto draw_bar = [&](std::vector<unsigned char>& nat, int state) {
    int filled = 4 * std::min(state, 10);
    for (int i = 0; i < SF_BAR_H * SF_BAR_W; i++) {
      int r = i / SF_BAR_W, c = i % SF_BAR_W;
      blend_px(nat, SF_BAR_X0 + c, SF_BAR_Y0 + r, t->colour_bar_bg, t->bar_alpha[r]);
      if (c < filled) blend_px(nat, SF_BAR_X0 + c, SF_BAR_Y0 + r, state == 11 ? t->colour_bar_kill : t->colour_bar_fg, t->bar_alpha[r]);
    }
  };
  std::vector<unsigned char> base(t->bg_nat, t->bg_nat + SF_NAT_H * SF_NAT_STRIDE);
  {
    std::vector<unsigned char> nat = base;
    draw_text0(nat);
    draw_bar(nat, 0);
    resample(nat, t->bg_obs);  // default observation: hexagons + "0000000" + empty bar
  }
  // which native rows do the text / bar output rows read?
  t->text_guard_row = t->yt_si[t->row_out1[SF_TEXT_Y0 + SF_TEXT_H - 1]][t->yt_cnt[t->row_out1[SF_TEXT_Y0 + SF_TEXT_H - 1]] - 1];
  t->bar_guard_row = t->yt_si[t->row_out0[SF_BAR_Y0]][0];
  if (t->row_out0[SF_BAR_Y0] * 84 != SF_BAR_CHUNK0 * 16 || t->row_out1[SF_BAR_Y0 + SF_BAR_H - 1] != 83) { snprintf(err, errcap, "bar output rows are not chunks 420..440"); return 1; }
  for (int st = 0; st < SF_BAR_STATES; st++) {
    std::vector<unsigned char> nat = base, obs(84 * 84);
    draw_text0(nat);
    draw_bar(nat, st);
    resample(nat, obs.data());
    memcpy(t->obs_bar[st], &obs[SF_BAR_CHUNK0 * 16], (SF_OBS_CHUNKS - SF_BAR_CHUNK0) * 16);
    for (int b = 0; b < SF_BAR_CHUNK0 * 16; b++) if (obs[b] != t->bg_obs[b]) { snprintf(err, errcap, "bar influences rows above chunk 420"); return 1; }
  }
  // sparse fortress lists + rects
  for (int k = 0; k < 36; k++) {
    int n = 0, x0 = 1 << 20, y0 = 1 << 20, x1 = -1, y1 = -1;
    for (int i = 0; i < SF_FORT_W * SF_FORT_W; i++) if (t->fort_alpha[k][i]) {
      if (n >= SF_FORT_LIST) { snprintf(err, errcap, "fortress sprite has more than %d lit pixels", SF_FORT_LIST); return 1; }
      int x = SF_FORT_X0 + i % SF_FORT_W, y = SF_FORT_Y0 + i / SF_FORT_W;
      t->fort_list_idx[k][n] = (unsigned short)i; t->fort_list_xy[k][n] = (unsigned short)(x | (y << 8)); t->fort_list_a[k][n] = t->fort_alpha[k][i]; n++;
      x0 = std::min(x0, x); y0 = std::min(y0, y); x1 = std::max(x1, x); y1 = std::max(y1, y);
    }
    t->fort_list_n[k] = n;
    t->fort_rect[k][0] = (unsigned char)x0; t->fort_rect[k][1] = (unsigned char)y0; t->fort_rect[k][2] = (unsigned char)x1; t->fort_rect[k][3] = (unsigned char)y1;
  }
  {
    int x0 = 1 << 20, y0 = 1 << 20, x1 = -1, y1 = -1;
    for (int i = 0; i < SF_EXP_W * SF_EXP_W; i++) if (t->fexp_alpha[0][i]) {
      int x = SF_FEXP_X0 + i % SF_EXP_W, y = SF_FEXP_Y0 + i / SF_EXP_W;
      x0 = std::min(x0, x); y0 = std::min(y0, y); x1 = std::max(x1, x); y1 = std::max(y1, y);
    }
    t->fort_rect[36][0] = (unsigned char)x0; t->fort_rect[36][1] = (unsigned char)y0; t->fort_rect[36][2] = (unsigned char)x1; t->fort_rect[36][3] = (unsigned char)y1;
  }
  // pre-resampled fortress states on the default base
  {
    std::vector<std::vector<unsigned char>> obs(SF_FORT_STATES, std::vector<unsigned char>(84 * 84));
    int c0 = SF_OBS_CHUNKS, c1 = -1;
    for (int st = 0; st < SF_FORT_STATES; st++) {
      std::vector<unsigned char> nat = base;
      if (st < 36) {
        for (int n = 0; n < t->fort_list_n[st]; n++) {
          int i = t->fort_list_idx[st][n];
          blend_px(nat, SF_FORT_X0 + i % SF_FORT_W, SF_FORT_Y0 + i / SF_FORT_W, t->colour_white, t->fort_list_a[st][n]);
        }
      } else {
        for (int i = 0; i < SF_EXP_W * SF_EXP_W; i++)
          for (int l = 0; l < t->fexp_layers && t->fexp_alpha[l][i]; l++)
            blend_px(nat, SF_FEXP_X0 + i % SF_EXP_W, SF_FEXP_Y0 + i / SF_EXP_W, t->fexp_colour[l][i], t->fexp_alpha[l][i]);
      }
      draw_text0(nat);
      draw_bar(nat, 0);
      resample(nat, obs[st].data());
      for (int c = 0; c < SF_OBS_CHUNKS; c++)
        if (memcmp(&obs[st][c * 16], &t->bg_obs[c * 16], 16)) { c0 = std::min(c0, c); c1 = std::max(c1, c); }
    }
    if (c1 < c0 || c1 - c0 + 1 > SF_FORT_CHUNKS || c1 >= SF_BAR_CHUNK0) { snprintf(err, errcap, "fortress chunk range %d..%d", c0, c1); return 1; }
    t->fort_chunk0 = c0; t->fort_nchunks = c1 - c0 + 1;
    for (int st = 0; st < SF_FORT_STATES; st++) memcpy(t->obs_fort[st], &obs[st][c0 * 16], (size_t)t->fort_nchunks * 16);
    for (int st = 0; st < SF_FORT_STATES; st++) {
      int n = 0;
      for (int c = 0; c < t->fort_nchunks; c++)
        if (memcmp(&t->obs_fort[st][c * 16], &t->bg_obs[(c0 + c) * 16], 16)) {
          if (n >= 64 || c > 254) { snprintf(err, errcap, "fortress state %d changes more than 64 chunks", st); return 1; }
          t->fort_sparse[st][n++] = (unsigned char)c;
        }
      t->fort_sparse_n[st] = (unsigned char)n;
      for (int k = n; k < 64; k++) t->fort_sparse[st][k] = 255;
    }
  }
  for (int a = 0; a < 360; a++) { t->hot.cs[a][0] = t->cos_deg[a]; t->hot.cs[a][1] = t->sin_deg[a]; }
  for (int h = 0; h < 2; h++)
    for (int k = 0; k < 6; k++) { t->hot.hex[h][k][0] = t->hex_px[h][k]; t->hot.hex[h][k][1] = t->hex_py[h][k]; t->hot.hex[h][k][2] = t->hex_nx[h][k]; t->hot.hex[h][k][3] = t->hex_ny[h][k]; }
  for (int k = 0; k < 8; k++) t->hot.atan2_oct[k] = t->atan2_oct[k];
  t->hot.ship_start_vx = t->ship_start_vx; t->hot.ship_start_vy = t->ship_start_vy;
  {
    const double radii[3] = {13.0, 23.0, 21.0};
    for (int k = 0; k < 3; k++) {
      const double r = radii[k];
      double x = r * r;
      while (sqrt(x) > r) x = nextafter(x, 0.0);                    // (never taken for these radii: r*r is exact)
      while (sqrt(nextafter(x, INFINITY)) <= r) x = nextafter(x, INFINITY);
      // exhaustive check of the claim around the threshold: 4096 doubles on each side
      double lo = x, hi = nextafter(x, INFINITY);
      for (int i = 0; i < 4096; i++) {
        if (!(sqrt(lo) <= r) || (sqrt(hi) <= r)) { snprintf(err, errcap, "touch threshold for r=%g is not a cut", r); return 1; }
        lo = nextafter(lo, 0.0); hi = nextafter(hi, INFINITY);
      }
      t->hot.touch2[k] = x;
    }
    t->hot.touch2[3] = 0.0;
  }
  return 0;
}

// ---- host-only composition of the static layers (CPU test hook, include/sf_b200.h) ----
#include "../../include/sf_b200.h"
extern "C" int sf_host_static_frame(int fortress_alive, int fortress_angle_deg, int points, int vulnerability, int kill_bar,
                                    uint8_t* h_native, uint8_t* h_bg_obs) {
  static SfTables* T = nullptr;
  if (!T) {
    T = new SfTables();
    char err[256];
    if (sf_build_tables(T, err, sizeof(err))) { delete T; T = nullptr; return SF_ERR_INVALID; }
  }
  if (!h_native || fortress_angle_deg < 0 || fortress_angle_deg >= 360 || fortress_angle_deg % 10) return SF_ERR_INVALID;
  std::vector<unsigned char> nat(T->bg_nat, T->bg_nat + SF_NAT_H * SF_NAT_STRIDE);
  auto px = [&](int x, int y) -> unsigned char& { return nat[y * SF_NAT_STRIDE + x]; };
  if (fortress_alive > 0) {
    const unsigned char* A = T->fort_alpha[fortress_angle_deg / 10];
    for (int i = 0; i < SF_FORT_W * SF_FORT_W; i++) if (A[i]) {
      unsigned char& d = px(SF_FORT_X0 + i % SF_FORT_W, SF_FORT_Y0 + i / SF_FORT_W);
      d = (unsigned char)sf_blend(d, T->colour_white, A[i]);
    }
  } else if (fortress_alive == 0) {
    for (int i = 0; i < SF_EXP_W * SF_EXP_W; i++)
      for (int l = 0; l < T->fexp_layers && T->fexp_alpha[l][i]; l++) {
        unsigned char& d = px(SF_FEXP_X0 + i % SF_EXP_W, SF_FEXP_Y0 + i / SF_EXP_W);
        d = (unsigned char)sf_blend(d, T->fexp_colour[l][i], T->fexp_alpha[l][i]);
      }
  }
  int pts = std::min(std::max(points, 0), 9999999);
  for (int i = 0; i < SF_TEXT_H * SF_TEXT_W && points >= 0; i++) {
    int slot = T->text_slot[i % SF_TEXT_W];
    if (slot >= 7) continue;
    int div = 1;
    for (int k = slot; k < 6; k++) div *= 10;
    unsigned a = T->text_alpha[(pts / div) % 10][i];
    if (a) { unsigned char& d = px(SF_TEXT_X0 + i % SF_TEXT_W, SF_TEXT_Y0 + i / SF_TEXT_W); d = (unsigned char)sf_blend(d, T->colour_text, a); }
  }
  int filled = 4 * std::min(vulnerability, 10);
  for (int i = 0; i < SF_BAR_H * SF_BAR_W && vulnerability >= 0; i++) {
    int r = i / SF_BAR_W, c = i % SF_BAR_W;
    unsigned char& d = px(SF_BAR_X0 + c, SF_BAR_Y0 + r);
    unsigned v = sf_blend(d, T->colour_bar_bg, T->bar_alpha[r]);
    if (c < filled) v = sf_blend(v, kill_bar ? T->colour_bar_kill : T->colour_bar_fg, T->bar_alpha[r]);
    d = (unsigned char)v;
  }
  for (int r = 0; r < SF_NAT_H; r++) memcpy(h_native + r * SF_NAT_W, &nat[r * SF_NAT_STRIDE], SF_NAT_W);
  if (h_bg_obs) memcpy(h_bg_obs, T->bg_obs, 84 * 84);
  return SF_OK;
}
