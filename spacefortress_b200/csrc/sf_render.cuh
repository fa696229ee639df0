// sf_render.cuh — warp-cooperative rasteriser: one warp draws one env's frame.
// Replaces drawGameStateScaled (draw.cpp:256-270), the RGBA2GRAY conversion (ssf_env.py:205, identity on grey
// input) and cv2.resize(INTER_AREA, 84x84) (rl/envs.py:29).
//
// Draw order and semantics follow draw.cpp:227-269:
//   black, 2 hexagons -> ship wireframe | ship explosion -> fortress wireframe | fortress explosion
//   -> missiles -> shells further than 21 from the fortress -> score digits -> vulnerability bar.
//
// Organisation (per env, all in one warp, everything in shared memory until the final stores):
//  * A persistent native tile (92x90 u8) holds the static background; every env composites its layers into
//    it, resamples only the touched rectangles, and restores them afterwards.
//  * Moving strokes (ship, missiles, shells; the ship explosion once per death) are scan-converted in
//    BATCHES: lanes build the stroked quads (fp64 CTM, 24.8 fixed point), then the (stroke,row,sub-row)
//    samples of the whole batch are flattened over the lanes; each sample evaluates its stroke's quads with
//    exact integer edge stepping (one "down" and one "up" edge are live per quad and sub-row), merges
//    overlapping spans (non-zero winding == union for equally oriented convex quads) and accumulates span
//    lengths into 16-bit cells. Regions are then blended in draw order.
//  * The ship explosion is identical for the 30 ticks a ship stays dead: it is rasterised once and kept in
//    a per-env 28x28 sprite cache (memo, not game state).
//  * Static layers (fortress sprite per sector angle, fortress explosion, score, vulnerability bar) come
//    from host-built tables. When no moving rectangle overlaps them their pre-resampled 16-byte OUTPUT
//    chunks are copied straight into the observation; otherwise they are blended into the tile.
//  * The observation is assembled from 441 16-byte chunks (background or static-layer tables) written with
//    coalesced 128-bit stores, then the resampled dirty pixels are patched in.
#pragma once
#include "sf_geom.h"
#include "sf_state.cuh"
#include "sf_tables.h"

#define SF_ACC_CELLS 1024    // 16-bit span-length cells per batch
#define SF_BATCH_QUADS 32
#define SF_BATCH_STROKES 32
#define SF_MAX_RECTS 32
#define SF_YBIAS 4096        // grid rows are stored biased so they fit an unsigned 16-bit field
#define SF_QUAD_IRREGULAR (1 << 20)  // quadrec.y flag: not a 2+2 edge split, test all four edges

// per-warp shared memory (14 KB -> 16 warps per SM)
struct __align__(16) SfWarpSmem {
  unsigned char nat[SF_NAT_H * SF_TILE_STRIDE];  // 8464 B persistent native tile
  unsigned short acc[SF_ACC_CELLS];              // 2048 B
  int4 edge[SF_BATCH_QUADS * 4];                 // 2048 B per quad: down0, down1, up0, up1 = {x_top, (ytop+bias)<<16 | dy, dx, magic}
  int2 quadrec[SF_BATCH_QUADS];                  //  256 B {ytopQ+bias, ybotQ+bias} (grid rows)
  int4 region[SF_BATCH_STROKES];                 //  512 B {x0, y0, w | h<<16, acc_off | colour<<16}
  int2 stroke[SF_BATCH_STROKES];                 //  256 B {region | quad0<<8 | nq<<16, first item}
  unsigned rect[SF_MAX_RECTS];                   //  128 B dirty rects x0 | y0<<8 | x1<<16 | y1<<24
  int nrect, nregion, nstroke, nitems;
  int acc_used, pad0, pad1, pad2;
};

// all kernels that render use the same dynamic shared array: one SfWarpSmem per warp. Helpers that are kept
// out of line re-derive their warp's slot from it, so the compiler still knows the address space.
extern __shared__ __align__(16) unsigned char sf_smem_raw[];
__device__ __forceinline__ SfWarpSmem& sf_my_smem() { return reinterpret_cast<SfWarpSmem*>(sf_smem_raw)[threadIdx.x >> 5]; }

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
__device__ __forceinline__ int sf_div_small(int a, int b, float inv_b) {  // a / b for 0 <= a < 2^20, exact
  return __float2int_rz(((float)a + 0.5f) * inv_b);
}

// once per warp at kernel start
__device__ __forceinline__ void sf_warp_smem_init(SfWarpSmem& W, const SfTables* T, int lane) {
  for (int k = lane; k < SF_ACC_CELLS / 2; k += 32) reinterpret_cast<unsigned*>(W.acc)[k] = 0u;
  for (int k = lane; k < SF_NAT_H * (SF_TILE_STRIDE / 4); k += 32) {
    int r = k / (SF_TILE_STRIDE / 4), c = k - r * (SF_TILE_STRIDE / 4);
    reinterpret_cast<unsigned*>(W.nat)[k] = __ldg(reinterpret_cast<const unsigned*>(T->bg_nat + r * SF_NAT_STRIDE) + c);
  }
  if (lane == 0) { W.nrect = 0; W.nregion = 0; W.nstroke = 0; W.nitems = 0; W.acc_used = 0; }
  __syncwarp();
}

__device__ __forceinline__ void sf_add_rect(SfWarpSmem& W, int lane, int x0, int y0, int x1, int y1) {
  if (lane == 0) {
    int n = W.nrect;
    unsigned r = (unsigned)x0 | ((unsigned)y0 << 8) | ((unsigned)x1 << 16) | ((unsigned)y1 << 24);
    if (n < SF_MAX_RECTS) { W.rect[n] = r; W.nrect = n + 1; }
    else {  // overflow: grow the last rectangle to the union (still correct, just more resampling)
      unsigned q = W.rect[SF_MAX_RECTS - 1];
      int qx0 = q & 255, qy0 = (q >> 8) & 255, qx1 = (q >> 16) & 255, qy1 = q >> 24;
      W.rect[SF_MAX_RECTS - 1] = (unsigned)min(qx0, x0) | ((unsigned)min(qy0, y0) << 8) | ((unsigned)max(qx1, x1) << 16) | ((unsigned)max(qy1, y1) << 24);
    }
  }
}

// ---- edge records --------------------------------------------------------------------------------------
// x(s) = x_top + floor((s - ytop) * dx / dy), exact, via a magic reciprocal M = floor((2^32-1)/dy) + 1 (valid while
// (dy*|dx|+dy)*dy < 2^32: always true for the strokes drawn here, which are at most ~30 px tall).
__device__ __forceinline__ int4 sf_make_edge(const SfTables* T, int xa, int ga, int xb, int gb) {  // ga < gb (grid rows)
  int dy = gb - ga;
  unsigned M = T->magic[min(dy, SF_MAGIC_N - 1)];  // device strokes are far shorter than 512 sub-rows (34 px)
  return make_int4(xa, ((ga + SF_YBIAS) << 16) | dy, xb - xa, (int)M);
}
__device__ __forceinline__ int sf_edge_x(int4 E, int sb) {  // sb = s + bias, ytop <= s < ytop + dy
  int m = sb - (int)((unsigned)E.y >> 16);
  int dy = E.y & 0xFFFF;
  unsigned adx = (unsigned)abs(E.z);
  unsigned n = (unsigned)m * adx + (E.z < 0 ? (unsigned)(dy - 1) : 0u);
  unsigned q = __umulhi(n, (unsigned)E.w);
  return E.z < 0 ? E.x - (int)q : E.x + (int)q;
}

// Build the 4 edge records of convex quad q (cyclic corners, consistent orientation) into slot `qi` and return
// its bounding box {ymin_g, ymax_g, xmin, xmax} (ymin_g >= ymax_g when the quad covers no sample).
// Edges whose grid rows increase along the traversal lie on one side, the others on the opposite side; a stroked
// segment is a parallelogram, so each side has at most two non-degenerate edges. Trapezoids (explosion arcs that
// straddle 0/180 degrees) can have three on one side: those keep all edges and the span tests every one.
__device__ __noinline__ int4 sf_store_quad_edges(const SfTables* T, int qi, int x0, int y0, int x1, int y1, int x2, int y2, int x3, int y3) {
  SfWarpSmem& W = sf_my_smem();
  const int g0 = sf_grid_y(y0), g1 = sf_grid_y(y1), g2 = sf_grid_y(y2), g3 = sf_grid_y(y3);
  const int4 none = make_int4(0, 0, 0, 0);  // dy 0: never live
  int4 E0 = none, E1 = none, E2 = none, E3 = none;
  int d0 = 0, d1 = 0, d2 = 0, d3 = 0;  // +1: rows increase along the traversal, -1: decrease, 0: degenerate
  if (g0 < g1) { E0 = sf_make_edge(T, x0, g0, x1, g1); d0 = 1; } else if (g0 > g1) { E0 = sf_make_edge(T, x1, g1, x0, g0); d0 = -1; }
  if (g1 < g2) { E1 = sf_make_edge(T, x1, g1, x2, g2); d1 = 1; } else if (g1 > g2) { E1 = sf_make_edge(T, x2, g2, x1, g1); d1 = -1; }
  if (g2 < g3) { E2 = sf_make_edge(T, x2, g2, x3, g3); d2 = 1; } else if (g2 > g3) { E2 = sf_make_edge(T, x3, g3, x2, g2); d2 = -1; }
  if (g3 < g0) { E3 = sf_make_edge(T, x3, g3, x0, g0); d3 = 1; } else if (g3 > g0) { E3 = sf_make_edge(T, x0, g0, x3, g3); d3 = -1; }
  const int nd = (d0 > 0) + (d1 > 0) + (d2 > 0) + (d3 > 0), nu = (d0 < 0) + (d1 < 0) + (d2 < 0) + (d3 < 0);
  const int gmin = min(min(g0, g1), min(g2, g3)), gmax = max(max(g0, g1), max(g2, g3));
  if (nd == 0 || nu == 0) { W.quadrec[qi] = make_int2(0, 0); return make_int4(1 << 30, -(1 << 30), 1 << 30, -(1 << 30)); }
  if (nd > 2 || nu > 2) {
    W.edge[qi * 4 + 0] = E0; W.edge[qi * 4 + 1] = E1; W.edge[qi * 4 + 2] = E2; W.edge[qi * 4 + 3] = E3;
    W.quadrec[qi] = make_int2(gmin + SF_YBIAS, (gmax + SF_YBIAS) | SF_QUAD_IRREGULAR);
  } else {
    // first / second edge of each direction in cyclic order
    int4 dnA = d0 > 0 ? E0 : d1 > 0 ? E1 : d2 > 0 ? E2 : E3;
    int4 dnB = d3 > 0 ? E3 : d2 > 0 ? E2 : d1 > 0 ? E1 : E0;
    int4 upA = d0 < 0 ? E0 : d1 < 0 ? E1 : d2 < 0 ? E2 : E3;
    int4 upB = d3 < 0 ? E3 : d2 < 0 ? E2 : d1 < 0 ? E1 : E0;
    // order each side top to bottom; a side with one edge has A == B (the split test selects B, same edge)
    if ((unsigned)dnB.y < (unsigned)dnA.y) { int4 t = dnA; dnA = dnB; dnB = t; }
    if ((unsigned)upB.y < (unsigned)upA.y) { int4 t = upA; upA = upB; upB = t; }
    W.edge[qi * 4 + 0] = dnA; W.edge[qi * 4 + 1] = dnB; W.edge[qi * 4 + 2] = upA; W.edge[qi * 4 + 3] = upB;
    W.quadrec[qi] = make_int2(gmin + SF_YBIAS, gmax + SF_YBIAS);
  }
  return make_int4(gmin, gmax, min(min(x0, x1), min(x2, x3)), max(max(x0, x1), max(x2, x3)));
}

// span of quad slot qi on biased grid row sb; returns false when the quad is not live there
__device__ __noinline__ void sf_quad_span_irregular(int qi, int sb, int& lo, int& hi) {
  const SfWarpSmem& W = sf_my_smem();
  lo = 1 << 30; hi = -(1 << 30);
#pragma unroll 1
  for (int k = 0; k < 4; k++) {
    int4 E = W.edge[qi * 4 + k];
    int m = sb - (int)((unsigned)E.y >> 16);
    if ((unsigned)m < (unsigned)(E.y & 0xFFFF)) { int x = sf_edge_x(E, sb); lo = min(lo, x); hi = max(hi, x); }
  }
}
__device__ __forceinline__ bool sf_quad_span(const SfWarpSmem& W, int qi, int sb, int& lo, int& hi) {
  int2 qr = W.quadrec[qi];
  if (qr.y & SF_QUAD_IRREGULAR) {
    if (sb < qr.x || sb >= (qr.y & ~SF_QUAD_IRREGULAR)) return false;
    sf_quad_span_irregular(qi, sb, lo, hi);
    return true;
  }
  if (sb < qr.x || sb >= qr.y) return false;
  int4 d1 = W.edge[qi * 4 + 1], u1 = W.edge[qi * 4 + 3];
  int4 d = (sb < (int)((unsigned)d1.y >> 16)) ? W.edge[qi * 4 + 0] : d1;
  int4 u = (sb < (int)((unsigned)u1.y >> 16)) ? W.edge[qi * 4 + 2] : u1;
  int xd = sf_edge_x(d, sb), xu = sf_edge_x(u, sb);
  lo = min(xd, xu); hi = max(xd, xu);
  return true;
}

// ---- batch machinery ---------------------------------------------------------------------------------------
__device__ __forceinline__ void sf_batch_begin(SfWarpSmem& W, int lane) {
  if (lane == 0) { W.nregion = 0; W.nstroke = 0; W.nitems = 0; W.acc_used = 0; }
  __syncwarp();
}

// floor division of a grid row by 15
__device__ __forceinline__ int sf_row_of(int g) { return (g >= 0) ? g / SF_GRID_Y : -((-g + SF_GRID_Y - 1) / SF_GRID_Y); }

// Open one region + one stroke per participating lane (`want`), in lane order, with a warp scan: region ids,
// accumulator offsets and first-item indices are exclusive prefixes. Returns the lane's region id or -1 (stroke
// off the surface, or the batch is full: cannot happen for the strokes drawn here, see the size notes above).
__device__ __forceinline__ int sf_open_regions(SfWarpSmem& W, int lane, bool want, int ymin_g, int ymax_g, int xmin, int xmax,
                                               unsigned colour, int quad0, int nq) {
  int cx0 = 0, py0 = 0, w = 0, h = 0;
  bool ok = want && ymin_g < ymax_g;
  if (ok) {
    py0 = max(sf_row_of(ymin_g), 0); int py1 = min(sf_row_of(ymax_g - 1), SF_NAT_H - 1);
    cx0 = max(xmin >> 8, 0); int cx1 = min((xmax - 1) >> 8, SF_NAT_W - 1);
    ok = py0 <= py1 && cx0 <= cx1;
    w = cx1 - cx0 + 1; h = py1 - py0 + 1;
  }
  int cells = ok ? w * h : 0;
  // inclusive scans of (count, cells) packed in one word: count < 64, cells < 2^20
  unsigned v = ok ? ((unsigned)cells | (1u << 24)) : 0u;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { unsigned t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
  int incl_cells = v & 0xFFFFFF, incl_cnt = v >> 24;
  if (ok && incl_cells > SF_ACC_CELLS) ok = false;  // pool full: drop (later strokes are dropped as well)
  unsigned okmask = __ballot_sync(0xffffffffu, ok);
  int rid = ok ? __popc(okmask & ((1u << lane) - 1u)) : -1;
  int items = ok ? h * SF_GRID_Y : 0;
  int iv = items;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, iv, o); if (lane >= o) iv += t; }
  if (ok) {
    W.region[rid] = make_int4(cx0, py0, w | (h << 16), (incl_cells - cells) | ((int)colour << 16));
    W.stroke[rid] = make_int2(rid | (quad0 << 8) | (nq << 16), iv - items);
  }
  if (lane == 31) { W.nregion = __popc(okmask); W.nstroke = __popc(okmask); W.nitems = iv; W.acc_used = incl_cells; }
  (void)incl_cnt;
  return rid;
}

// Accumulate span lengths of every stroke of the batch. Samples (stroke, row, sub-row) are flattened over lanes.
// Within a stroke, samples are ordered in blocks of 4 pixel rows (row fastest, then sub-row): one pass of 32 lanes
// then covers ~4 rows x 8 sub-rows, so (a) same-cell atomic conflicts stay low and (b) quads that do not reach
// those rows are skipped with a warp-uniform test.
__device__ __noinline__ void sf_batch_accumulate() {
  SfWarpSmem& W = sf_my_smem();
  const int lane = threadIdx.x & 31;
  __syncwarp();
  const int nitems = W.nitems, ns = W.nstroke;
  int s_first = 0;  // stroke containing the first sample of the pass (warp uniform)
#pragma unroll 1
  for (int it0 = 0; it0 < nitems; it0 += 32) {
    const int it = it0 + lane;
    const bool valid = it < nitems;
    while (s_first + 1 < ns && W.stroke[s_first + 1].y <= it0) s_first++;
    int si = s_first;
    while (valid && si + 1 < ns && W.stroke[si + 1].y <= it) si++;
    const int2 S = W.stroke[si];
    const int4 R = W.region[S.x & 255];
    const int quad0 = (S.x >> 8) & 255, nq = (S.x >> 16) & 255;
    const int w = R.z & 0xFFFF, h = (R.z >> 16) & 0xFFFF;
    // sample index -> (row, sub): full blocks of 4 rows, then the remaining 1..3 rows
    const int li = valid ? it - S.y : 0;
    const int nfull = h >> 2, rem = h & 3;
    int r, sub;
    if (li < nfull * 60) {
      int blk = sf_div_small(li, 60, 1.0f / 60.0f), t = li - blk * 60;
      sub = t >> 2; r = blk * 4 + (t & 3);
    } else {
      int t = li - nfull * 60;
      sub = rem == 3 ? sf_div_small(t, 3, 1.0f / 3.0f) : (rem == 2 ? t >> 1 : t);
      r = nfull * 4 + t - sub * rem;
    }
    const int sb = (R.y + r) * SF_GRID_Y + sub + SF_YBIAS;
    const int xlo = R.x << 8, xhi = (R.x + w) << 8;
    const int cell0 = (R.w & 0xFFFF) + r * w;
    // spans of the stroke's quads, kept sorted by start in four registers (insertion keeps the loops rolled so
    // the whole body stays small enough for the instruction cache)
    unsigned k0 = 0xFFFFFFFFu, k1 = 0xFFFFFFFFu, k2 = 0xFFFFFFFFu, k3 = 0xFFFFFFFFu;
#pragma unroll 1
    for (int j = 0; j < 4; j++) {
      bool live = false;
      if (valid && j < nq) {
        int2 qr = W.quadrec[quad0 + j];
        live = sb >= qr.x && sb < (qr.y & ~SF_QUAD_IRREGULAR);
      }
      if (!__any_sync(0xffffffffu, live)) continue;
      unsigned key = 0xFFFFFFFFu;
      if (live) {
        int lo, hi;
        sf_quad_span(W, quad0 + j, sb, lo, hi);
        lo = max(lo, xlo); hi = min(hi, xhi);
        if (lo < hi) key = ((unsigned)(lo - xlo) << 16) | (unsigned)(hi - xlo);
      }
      unsigned t = key, m;
      m = min(k0, t); t = max(k0, t); k0 = m;
      m = min(k1, t); t = max(k1, t); k1 = m;
      m = min(k2, t); t = max(k2, t); k2 = m;
      k3 = min(k3, t);
    }
    // emit each span minus the union of its predecessors; 16-bit cells, 32-bit atomics: a cell never exceeds
    // 15*256, so the two halves of a word cannot carry into each other
    unsigned* acc32 = reinterpret_cast<unsigned*>(W.acc);
    int reach = 0;
#pragma unroll 1
    while (__any_sync(0xffffffffu, k0 != 0xFFFFFFFFu)) {
      if (k0 != 0xFFFFFFFFu) {
        int a = max((int)(k0 >> 16), reach), b = (int)(k0 & 0xFFFFu);
        reach = max(reach, b);
        if (a < b) {
          // first cell (partial), last cell (partial), full cells in between (only for near-horizontal spans)
          int c1 = a >> 8, c2 = (b - 1) >> 8;
          int ci = cell0 + c1;
          int len1 = min(b, (c1 + 1) << 8) - a;
          atomicAdd(&acc32[ci >> 1], (unsigned)len1 << ((ci & 1) << 4));
          if (c2 > c1) {
            int cj = cell0 + c2;
            atomicAdd(&acc32[cj >> 1], (unsigned)(b - (c2 << 8)) << ((cj & 1) << 4));
            for (int c = c1 + 1; c < c2; c++) { int cm = cell0 + c; atomicAdd(&acc32[cm >> 1], 256u << ((cm & 1) << 4)); }
          }
        }
      }
      k0 = k1; k1 = k2; k2 = k3; k3 = 0xFFFFFFFFu;
    }
  }
  __syncwarp();
}

// Blend one region into the tile (and zero its cells), record its dirty rectangle.
__device__ __noinline__ void sf_region_blend(int region_id) {
  SfWarpSmem& W = sf_my_smem();
  const int lane = threadIdx.x & 31;
  int4 R = W.region[region_id];
  int w = R.z & 0xFFFF, h = (R.z >> 16) & 0xFFFF;
  unsigned colour = (unsigned)R.w >> 16;
  unsigned short* cells = W.acc + (R.w & 0xFFFF);
  float inv_w = 1.0f / (float)w;
  for (int idx = lane; idx < w * h; idx += 32) {
    unsigned L = cells[idx];
    if (L) {
      cells[idx] = 0;
      int r = sf_div_small(idx, w, inv_w), c = idx - r * w;
      unsigned char* px = &W.nat[(R.y + r) * SF_TILE_STRIDE + R.x + c];
      *px = (unsigned char)sf_blend(*px, colour, sf_len_to_alpha(L));
    }
  }
  sf_add_rect(W, lane, R.x, R.y, R.x + w - 1, R.y + h - 1);
  __syncwarp();
}

// ---- wireframe strokes (R3 drawWireFrame, draw.cpp:82-100) --------------------------------------------------------
// Geometry for up to 8 strokes at once: lane = 4*slot + line. kind: 0 ship, 1 missile, 2 shell, -1 none.
// Every lane passes the description of ITS slot's stroke. Returns (by all lanes) the region id of the lane's slot
// or -1 when that stroke is invisible / did not fit (the caller retries it in the next batch).
__device__ __forceinline__ int sf_wire_geometry(SfWarpSmem& W, int lane, const SfTables* T, int kind, double px, double py, int angle) {
  const int slot = lane >> 2, line = lane & 3;
  int ymin_g = 1 << 30, ymax_g = -(1 << 30), xmin = 1 << 30, xmax = -(1 << 30);
  bool has = false;
  if (kind >= 0) {
    // quick cull: every model fits in a 37-unit radius (7.4 px) around its origin
    double dxv = SF_DADD(SF_DMUL(px, SF_CTM_SCALE), SF_CTM_X0), dyv = SF_DADD(SF_DMUL(py, SF_CTM_SCALE), SF_CTM_Y0);
    bool visible = !(dxv < -9.0 || dxv > SF_NAT_W + 9.0 || dyv < -9.0 || dyv > SF_NAT_H + 9.0);
    if (visible && line < T->wf_nlines[kind]) {
      SfWireXf m = sf_wire_xf(px, py, T->cos_deg[angle], T->sin_deg[angle]);
      const double* L = T->wf_line[kind][line];
      SfPt a = sf_xform_wire(m, L[0], L[1]), b = sf_xform_wire(m, L[2], L[3]);
      SfQuad q;
      if (sf_stroke_quad(a, b, q)) {
        int4 bb = sf_store_quad_edges(T, slot * 4 + line, q.p[0].x, q.p[0].y, q.p[1].x, q.p[1].y, q.p[2].x, q.p[2].y, q.p[3].x, q.p[3].y);
        ymin_g = bb.x; ymax_g = bb.y; xmin = bb.z; xmax = bb.w; has = true;
      }
    }
  }
  if (!has) W.quadrec[slot * 4 + line] = make_int2(0, 0);
  // bounding box of the slot's stroke: reduce over its 4 lanes
#pragma unroll
  for (int o = 1; o <= 2; o <<= 1) {
    ymin_g = min(ymin_g, __shfl_xor_sync(0xffffffffu, ymin_g, o)); ymax_g = max(ymax_g, __shfl_xor_sync(0xffffffffu, ymax_g, o));
    xmin = min(xmin, __shfl_xor_sync(0xffffffffu, xmin, o)); xmax = max(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
  }
  // one region + stroke per slot, opened in slot order by the slot's first lane
  int region_id = sf_open_regions(W, lane, line == 0 && kind >= 0, ymin_g, ymax_g, xmin, xmax, T->colour_white, slot * 4, 4);
  return __shfl_sync(0xffffffffu, region_id, lane & ~3);
}

// ---- ship explosion (R5 drawExplosion, draw.cpp:116-145): 84 arcs (one stroke each) + the r=7 circle -------------
__device__ __noinline__ void sf_explosion_raster(const SfTables* T, double px, double py) {
  SfWarpSmem& W = sf_my_smem();
  const int lane = threadIdx.x & 31;
  SfPt c = sf_xform_base(px, py);
  // arcs in batches of 32 strokes (one quad each); the last batch is the circle: 16 abutting quads = 4 strokes of
  // 4 quads sharing one region
#pragma unroll 1
  for (int s0 = 0; s0 < SF_EXP_STROKES - 1 + 32; s0 += 32) {
    const bool circle = s0 >= SF_EXP_STROKES - 1;
    sf_batch_begin(W, lane);
    int s = circle ? SF_EXP_STROKES - 1 + lane : s0 + lane;
    int ymin_g = 1 << 30, ymax_g = -(1 << 30), xmin = 1 << 30, xmax = -(1 << 30);
    bool mine = circle ? lane < 16 : s < SF_EXP_STROKES - 1;
    if (mine) {
      const short* o = T->exp_quad[s];
      int4 bb = sf_store_quad_edges(T, lane, c.x + o[0], c.y + o[1], c.x + o[2], c.y + o[3], c.x + o[4], c.y + o[5], c.x + o[6], c.y + o[7]);
      ymin_g = bb.x; ymax_g = bb.y; xmin = bb.z; xmax = bb.w;
    } else W.quadrec[lane] = make_int2(0, 0);
    if (!circle) {
      sf_open_regions(W, lane, mine, ymin_g, ymax_g, xmin, xmax, mine ? T->exp_colour[s] : 0u, lane, 1);
    } else {
      ymin_g = sf_warp_min(ymin_g); ymax_g = sf_warp_max(ymax_g); xmin = sf_warp_min(xmin); xmax = sf_warp_max(xmax);
      int rid = sf_open_regions(W, lane, lane == 0, ymin_g, ymax_g, xmin, xmax, T->exp_colour[SF_EXP_STROKES - 1], 0, 4);
      rid = __shfl_sync(0xffffffffu, rid, 0);
      __syncwarp();
      if (rid >= 0 && lane >= 1 && lane < 4) {  // three more strokes over the same region
        int h = (W.region[0].z >> 16) & 0xFFFF;
        W.stroke[lane] = make_int2(0 | ((lane * 4) << 8) | (4 << 16), lane * h * SF_GRID_Y);
      }
      __syncwarp();
      if (rid >= 0 && lane == 0) { int h = (W.region[0].z >> 16) & 0xFFFF; W.nstroke = 4; W.nitems = 4 * h * SF_GRID_Y; }
    }
    sf_batch_accumulate();
    const int nreg = W.nregion;
#pragma unroll 1
    for (int r = 0; r < nreg; r++) sf_region_blend(r);  // regions are in stroke order
  }
}

// one output pixel of cv2 INTER_AREA (float accumulation in table order, round-half-even)
__device__ __forceinline__ unsigned char sf_resample(const SfWarpSmem& W, const SfTables* T, int i, int j) {
  SfTap ty = T->ytap[i], tx = T->xtap[j];
  const unsigned char* S = &W.nat[ty.si * SF_TILE_STRIDE + tx.si];
  float sum = 0.f;
#pragma unroll
  for (int ky = 0; ky < SF_MAX_TAPS; ky++) {
    if (ky < ty.cnt) {
      float buf = 0.f;
#pragma unroll
      for (int kx = 0; kx < 2; kx++)  // the x table never has more than 2 taps (scale 15/14)
        if (kx < tx.cnt) buf = __fadd_rn(buf, __fmul_rn((float)S[ky * SF_TILE_STRIDE + kx], tx.a[kx]));
      float v = __fmul_rn(ty.a[ky], buf);
      sum = ky == 0 ? v : __fadd_rn(sum, v);
    }
  }
  return (unsigned char)__float2int_rn(sum);
}

// Ship explosion layer: rasterised once per death, then replayed from the per-env sprite cache.
__device__ __noinline__ void sf_ship_explosion(const SfTables* T, unsigned char* expc, int4* q0, int env, unsigned core, double px, double py) {
  SfWarpSmem& W = sf_my_smem();
  const int lane = threadIdx.x & 31;
  SfPt c = sf_xform_base(px, py);
  int bx0 = (c.x >> 8) - 13, by0 = (c.y >> 8) - 13;
  unsigned char* cache = expc + (size_t)env * (SF_EXP_W * SF_EXP_W);
  if (!(core & SF_CORE_EXP_CACHED)) {
    sf_explosion_raster(T, px, py);
#pragma unroll 1
    for (int idx = lane; idx < SF_EXP_W * SF_EXP_W; idx += 32) {
      int r = idx / SF_EXP_W, cc = idx - r * SF_EXP_W, x = bx0 + cc, y = by0 + r;
      if (x >= 0 && x < SF_NAT_W && y >= 0 && y < SF_NAT_H) cache[idx] = W.nat[y * SF_TILE_STRIDE + x];
    }
    if (lane == 0) q0[env].x = (int)(core | SF_CORE_EXP_CACHED);
    if (lane == 0) W.nrect = 0;  // the arcs recorded up to 85 small rectangles: replace them by the box
  } else {
#pragma unroll 1
    for (int idx = lane; idx < SF_EXP_W * SF_EXP_W; idx += 32) {
      int r = idx / SF_EXP_W, cc = idx - r * SF_EXP_W, x = bx0 + cc, y = by0 + r;
      if (x >= 0 && x < SF_NAT_W && y >= 0 && y < SF_NAT_H) W.nat[y * SF_TILE_STRIDE + x] = cache[idx];
    }
  }
  int x0 = max(bx0, 0), y0 = max(by0, 0), x1 = min(bx0 + SF_EXP_W - 1, SF_NAT_W - 1), y1 = min(by0 + SF_EXP_W - 1, SF_NAT_H - 1);
  if (x0 <= x1 && y0 <= y1) sf_add_rect(W, lane, x0, y0, x1, y1);
  __syncwarp();
}

// Fortress layer blended into the tile (general path: something that moves overlaps it, or native output).
__device__ __noinline__ void sf_fortress_general(const SfTables* T, int st) {
  SfWarpSmem& W = sf_my_smem();
  const int lane = threadIdx.x & 31;
  const unsigned char* fr = T->fort_rect[st];
  if (st < 36) {
    const int n = T->fort_list_n[st];
#pragma unroll 1
    for (int k = lane; k < n; k += 32) {
      int idx = T->fort_list_idx[st][k];
      int r = idx / SF_FORT_W, c = idx - r * SF_FORT_W;
      unsigned char* px = &W.nat[(SF_FORT_Y0 + r) * SF_TILE_STRIDE + SF_FORT_X0 + c];
      *px = (unsigned char)sf_blend(*px, T->colour_white, T->fort_list_a[st][k]);
    }
  } else {
#pragma unroll 1
    for (int idx = lane; idx < SF_EXP_W * SF_EXP_W; idx += 32) {
      unsigned a0 = T->fexp_alpha[0][idx];
      if (a0) {
        int r = idx / SF_EXP_W, c = idx - r * SF_EXP_W;
        unsigned char* px = &W.nat[(SF_FEXP_Y0 + r) * SF_TILE_STRIDE + SF_FEXP_X0 + c];
        unsigned v = sf_blend(*px, T->fexp_colour[0][idx], a0);
        for (int l = 1; l < T->fexp_layers; l++) {
          unsigned a = T->fexp_alpha[l][idx];
          if (!a) break;
          v = sf_blend(v, T->fexp_colour[l][idx], a);
        }
        *px = (unsigned char)v;
      }
    }
  }
  sf_add_rect(W, lane, fr[0], fr[1], fr[2], fr[3]);
  __syncwarp();
}

// Score digits (draw.cpp:160-173,267) and vulnerability bar (draw.cpp:207-225,268) blended into the tile.
__device__ __noinline__ void sf_text_general(const SfTables* T, int pts) {
  SfWarpSmem& W = sf_my_smem();
  const int lane = threadIdx.x & 31;
#pragma unroll 1
  for (int idx = lane; idx < SF_TEXT_H * SF_TEXT_W; idx += 32) {
    int r = idx / SF_TEXT_W, c = idx - r * SF_TEXT_W;
    int slot = T->text_slot[c];
    if (slot < 7) {
      int div = 1;
      for (int k = slot; k < 6; k++) div *= 10;
      unsigned a = T->text_alpha[(pts / div) % 10][idx];
      if (a) {
        unsigned char* px = &W.nat[(SF_TEXT_Y0 + r) * SF_TILE_STRIDE + SF_TEXT_X0 + c];
        *px = (unsigned char)sf_blend(*px, T->colour_text, a);
      }
    }
  }
  sf_add_rect(W, lane, SF_TEXT_X0, SF_TEXT_Y0, SF_TEXT_X0 + SF_TEXT_W - 1, SF_TEXT_Y0 + SF_TEXT_H - 1);
}
__device__ __noinline__ void sf_bar_general(const SfTables* T, int vuln, bool kill_bar) {
  SfWarpSmem& W = sf_my_smem();
  const int lane = threadIdx.x & 31;
  int filled = 4 * min(vuln, 10);  // 20 user units per step = 4 px
  unsigned fg = kill_bar ? T->colour_bar_kill : T->colour_bar_fg;
#pragma unroll 1
  for (int idx = lane; idx < SF_BAR_H * SF_BAR_W; idx += 32) {
    int r = idx / SF_BAR_W, c = idx - r * SF_BAR_W;
    unsigned a = T->bar_alpha[r];
    unsigned char* px = &W.nat[(SF_BAR_Y0 + r) * SF_TILE_STRIDE + SF_BAR_X0 + c];
    unsigned v = sf_blend(*px, T->colour_bar_bg, a);
    if (c < filled) v = sf_blend(v, fg, a);
    *px = (unsigned char)v;
  }
  sf_add_rect(W, lane, SF_BAR_X0, SF_BAR_Y0, SF_BAR_X0 + SF_BAR_W - 1, SF_BAR_Y0 + SF_BAR_H - 1);
}
__device__ __noinline__ void sf_native_out(unsigned char* __restrict__ nat_out) {
  SfWarpSmem& W = sf_my_smem();
  const int lane = threadIdx.x & 31;
#pragma unroll 1
  for (int idx = lane; idx < SF_NAT_H * SF_NAT_W / 2; idx += 32) {
    int r = idx / (SF_NAT_W / 2), c = (idx - r * (SF_NAT_W / 2)) * 2;
    *reinterpret_cast<uchar2*>(&nat_out[r * SF_NAT_W + c]) = *reinterpret_cast<const uchar2*>(&W.nat[r * SF_TILE_STRIDE + c]);
  }
}

struct SfRenderIn {  // warp-uniform view of one env
  int env;
  unsigned core, pmask;
  double px, py;
  int points_i, vuln;
  bool kill_bar;  // vuln > 10 && vulnerability timer < 250 (draw.cpp:268)
};

__device__ __forceinline__ bool sf_regions_touch(const SfWarpSmem& W, int nreg, int x0, int y0, int x1, int y1) {
  bool hit = false;
  for (int q = 0; q < nreg; q++) {
    int4 R = W.region[q];
    int rx1 = R.x + (R.z & 0xFFFF) - 1, ry1 = R.y + ((R.z >> 16) & 0xFFFF) - 1;
    hit |= !(rx1 < x0 || R.x > x1 || ry1 < y0 || R.y > y1);
  }
  return hit;
}
__device__ __forceinline__ bool sf_rects_touch(const SfWarpSmem& W, int nr, int x0, int y0, int x1, int y1) {
  bool hit = false;
  for (int q = 0; q < nr; q++) {
    unsigned r = W.rect[q];
    int rx0 = r & 255, ry0 = (r >> 8) & 255, rx1 = (r >> 16) & 255, ry1 = r >> 24;
    hit |= !(rx1 < x0 || rx0 > x1 || ry1 < y0 || ry0 > y1);
  }
  return hit;
}

// Draw env `in` and write its observation. obs84: 84*84 bytes (or NULL), nat_out: 92*90 bytes (or NULL).
__device__ __forceinline__ void sf_render_env(const SfDev& D, SfWarpSmem& W, int lane, const SfRenderIn& in,
                                              unsigned char* __restrict__ obs84, unsigned char* __restrict__ nat_out) {
  const SfTables* T = D.tab;
  const int np = D.n_pad;
  const bool ship_alive = in.core & SF_CORE_SHIP_ALIVE, fort_alive = in.core & SF_CORE_FORT_ALIVE;
  if (lane == 0) W.nrect = 0;

  // ---- stroke list: [ship] + live missiles (slot order) + visible shells (slot order) ----
  const unsigned mm = in.pmask & SF_PMASK_MISSILES;
  unsigned sm = 0;
  {
    bool vis = false;
    if (lane < SF_DEV_SHELLS && ((in.pmask >> (SF_PMASK_SHELL_SHIFT + lane)) & 1u)) {
      double2 p = D.spos[(size_t)lane * np + in.env];
      double dx = SF_DSUB(p.x, SF_FORT_X), dy = SF_DSUB(p.y, SF_FORT_Y);
      vis = SF_DSQRT(SF_DADD(SF_DMUL(dx, dx), SF_DMUL(dy, dy))) > 21.0;  // quirk Q9, draw.cpp:249-250
    }
    sm = __ballot_sync(0xffffffffu, vis) & 0xFu;
  }
  const int n_ship = ship_alive ? 1 : 0, n_mis = __popc(mm), n_strokes = n_ship + n_mis + __popc(sm);

  // ---- ship explosion: memoised sprite (draw.cpp:235-237) ----
  __syncwarp();
  if (!ship_alive) sf_ship_explosion(T, D.expc, D.q0, in.env, in.core, in.px, in.py);

  // ---- moving wireframes in batches of 8 strokes; the fortress layer goes in after the ship ----
  bool fortress_done = false;
  // fast path: nothing that moves comes near the sprite -> its pre-resampled output chunks are used and the
  // tile is left alone. Decided once every moving region is known (explosion box + first batch; with more
  // than 8 strokes the general path is taken).
  const int fst_ = fort_alive ? (int)((in.core >> SF_CORE_FANG_SHIFT) & 63u) : 36;
  bool fortress_general = nat_out != nullptr || n_strokes > 8 ||
                          sf_rects_touch(W, W.nrect, T->fort_rect[fst_][0] - 2, T->fort_rect[fst_][1] - 2, T->fort_rect[fst_][2] + 2, T->fort_rect[fst_][3] + 2);
  auto fortress_layer = [&]() { if (fortress_general) sf_fortress_general(T, fst_); };

  for (int s0 = 0; s0 < n_strokes || !fortress_done; s0 += 8) {
    int region_id = -1;
    if (s0 < n_strokes) {
      sf_batch_begin(W, lane);
      const int si = s0 + (lane >> 2);
      int kind = -1, angle = 0;
      double qx = 0, qy = 0;
      if (si < n_strokes) {
        if (si < n_ship) { kind = 0; qx = in.px; qy = in.py; angle = (int)(in.core & SF_CORE_ANGLE_MASK); }
        else if (si < n_ship + n_mis) {
          int slot = __fns(mm, 0, si - n_ship + 1);
          double2 p = D.mpos[(size_t)slot * np + in.env];
          kind = 1; qx = p.x; qy = p.y; angle = D.mang[(size_t)slot * np + in.env];
        } else {
          int slot = __fns(sm, 0, si - n_ship - n_mis + 1);
          double2 p = D.spos[(size_t)slot * np + in.env];
          kind = 2; qx = p.x; qy = p.y;
          angle = __double2int_rz(D.sang[(size_t)slot * np + in.env]);  // `int angle` truncation, quirk Q10
          if (angle >= 360) angle -= 360;
        }
      }
      region_id = sf_wire_geometry(W, lane, T, kind, qx, qy, angle);
      if (s0 == 0 && !fortress_general)
        fortress_general = sf_regions_touch(W, W.nregion, T->fort_rect[fst_][0] - 2, T->fort_rect[fst_][1] - 2, T->fort_rect[fst_][2] + 2, T->fort_rect[fst_][3] + 2);
      sf_batch_accumulate();
    }
    // blend in draw order: ship first, then the fortress layer, then projectiles
    for (int slot = 0; slot < 8; slot++) {
      int si = s0 + slot;
      if (!fortress_done && si >= n_ship) { fortress_layer(); fortress_done = true; }
      if (si >= n_strokes) break;
      int rid = __shfl_sync(0xffffffffu, region_id, slot * 4);
      if (rid >= 0) sf_region_blend(rid);
    }
  }

  // ---- score digits (draw.cpp:160-173,267): "%07d" of (int)mPoints ----
  const int nmoving = W.nrect;
  bool text_general, bar_general;
  {
    const int pts = min(max(in.points_i, 0), 9999999);
    bool moving_above = false, moving_below = false;
    for (int q = 0; q < nmoving; q++) {
      unsigned r = W.rect[q];
      moving_above |= (int)((r >> 8) & 255) <= T->text_guard_row;
      moving_below |= (int)(r >> 24) >= T->bar_guard_row;
    }
    text_general = nat_out != nullptr || pts != 0 || moving_above;
    bar_general = nat_out != nullptr || moving_below;
    if (text_general) sf_text_general(T, pts);
    if (bar_general) sf_bar_general(T, in.vuln, in.kill_bar);
  }
  __syncwarp();

  // ---- native output (SSF_Env.step returns the 92x90 frame): every layer went through the tile ----
  if (nat_out) sf_native_out(nat_out);

  // ---- 84x84 observation: 441 chunks from the static tables, then the resampled dirty rectangles ----
  if (obs84) {
    const int fst = fort_alive ? (int)((in.core >> SF_CORE_FANG_SHIFT) & 63u) : 36;
    const int bst = in.kill_bar ? 11 : min(in.vuln, 10);
    const int fc0 = T->fort_chunk0, fc1 = fc0 + T->fort_nchunks;
    int4* g = reinterpret_cast<int4*>(obs84);
    const int4* bg = reinterpret_cast<const int4*>(T->bg_obs);
    const int4* ft = reinterpret_cast<const int4*>(T->obs_fort[fst]);
    const int4* bt = reinterpret_cast<const int4*>(T->obs_bar[bst]);
    // three ranges with a fixed source each: [0, fc0) background, [fc0, fc1) fortress table (fast path),
    // [fc1, 420) background, [420, 441) bar table (fast path)
    const int4* fsrc = fortress_general ? bg + fc0 : ft;
    const int4* bsrc = bar_general ? bg + SF_BAR_CHUNK0 : bt;
#pragma unroll 1
    for (int k = lane; k < fc0; k += 32) g[k] = __ldg(&bg[k]);
#pragma unroll 1
    for (int k = fc0 + lane; k < fc1; k += 32) g[k] = __ldg(&fsrc[k - fc0]);
#pragma unroll 1
    for (int k = fc1 + lane; k < SF_BAR_CHUNK0; k += 32) g[k] = __ldg(&bg[k]);
    if (lane < SF_OBS_CHUNKS - SF_BAR_CHUNK0) g[SF_BAR_CHUNK0 + lane] = __ldg(&bsrc[lane]);
    __syncwarp();  // orders the chunk stores before the byte patches below (same warp)
    const int nr = W.nrect;
    for (int q = 0; q < nr; q++) {
      unsigned R = W.rect[q];
      int x0 = R & 255, y0 = (R >> 8) & 255, x1 = (R >> 16) & 255, y1 = R >> 24;
      int j0 = T->col_out0[x0], j1 = T->col_out1[x1], i0 = T->row_out0[y0], i1 = T->row_out1[y1];
      int ow = j1 - j0 + 1, cnt = ow * (i1 - i0 + 1);
      float inv_ow = 1.0f / (float)ow;
      for (int idx = lane; idx < cnt; idx += 32) {
        int r = sf_div_small(idx, ow, inv_ow), c = idx - r * ow;
        obs84[(i0 + r) * 84 + j0 + c] = sf_resample(W, T, i0 + r, j0 + c);
      }
    }
  }
  __syncwarp();

  // ---- restore the touched rectangles of the persistent tile ----
  {
    const int nr = W.nrect;
    for (int q = 0; q < nr; q++) {
      unsigned R = W.rect[q];
      int x0 = R & 255, y0 = (R >> 8) & 255, x1 = (R >> 16) & 255, y1 = R >> 24;
      int w = x1 - x0 + 1, cnt = w * (y1 - y0 + 1);
      float inv_w = 1.0f / (float)w;
      for (int idx = lane; idx < cnt; idx += 32) {
        int r = sf_div_small(idx, w, inv_w), c = idx - r * w;
        W.nat[(y0 + r) * SF_TILE_STRIDE + x0 + c] = T->bg_nat[(y0 + r) * SF_NAT_STRIDE + x0 + c];
      }
    }
  }
  __syncwarp();
}
