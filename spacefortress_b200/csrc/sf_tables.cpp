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

int sf_build_tables(SfTables* t, char* err, int errcap) {
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

  // ---- score digits: 7-segment face on the metrics of a 30-unit monospace bold font ----
  {
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

  // ---- bg_obs = INTER_AREA(bg_nat) with the same float operation order as the device epilogue ----
  for (int i = 0; i < 84; i++)
    for (int j = 0; j < 84; j++) {
      float sum = 0.f;
      for (int ky = 0; ky < t->yt_cnt[i]; ky++) {
        float buf = 0.f;
        const unsigned char* S = &t->bg_nat[t->yt_si[i][ky] * SF_NAT_STRIDE];
        for (int kx = 0; kx < t->xt_cnt[j]; kx++) buf += (float)S[t->xt_si[j][kx]] * t->xt_a[j][kx];
        sum = (ky == 0) ? t->yt_a[i][ky] * buf : sum + t->yt_a[i][ky] * buf;
      }
      t->bg_obs[i * 84 + j] = (unsigned char)lrintf(sum);
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
