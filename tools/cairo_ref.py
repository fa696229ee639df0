"""Real-cairo rendition of the reference's frame, for machines that HAVE cairo (pycairo or cairocffi) and a "monospace"
font: a call-for-call Python transcription of drawGameStateScaled / drawJustGameStuff / drawScore / drawVlner
(python/spacefortress/src/draw.cpp:82-145,147-173,207-270) on a 90x92 RGB24 image surface with the gym env's parameters
(ssf_env.py:50,164: viewport (130,80,450,460), scale .2, line width 3, grayscale). Used by tools/compare_cairo.py and
tools/dump_cairo_glyphs.py; not imported by the package, the tests or the bench (cairo is absent from the build image,
which is why frame parity against real cairo is unpinned: DESIGN.md §5)."""
import json, math, os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def import_cairo():
    try:
        import cairo  # pycairo
        return cairo
    except Exception:
        try:
            import cairocffi as cairo
            return cairo
        except Exception:
            return None


# wireframe.cpp:11-67 — points and lines of the four models
WIREFRAMES = {
    "missile": ([(0, 0), (-25, 0), (-5, 5), (-5, -5)], [(0, 1), (0, 2), (0, 3)]),
    "shell": ([(-8, 0), (0, -6), (16, 0), (0, 6), (-8, 0)], [(0, 1), (1, 2), (2, 3), (3, 0)]),
    "ship": ([(-18, 0), (18, 0), (0, 0), (-18, 18), (-18, -18)], [(0, 1), (3, 2), (2, 4)]),
    "fortress": ([(0, 0), (36, 0), (18, -18), (0, -18), (18, 18), (0, 18)], [(0, 1), (3, 2), (2, 4), (4, 5)]),
}
KA = json.load(open(os.path.join(ROOT, "tests", "golden", "known_answers.json")))  # hexagon vertices (hexagon.cpp:13-35)


def _wireframe(ctx, name, x, y, angle, ls, grey):          # draw.cpp:82-100 (`int angle`)
    pts, lines = WIREFRAMES[name]
    ctx.save()
    ctx.translate(x, y)
    ctx.rotate(int(angle) * math.pi / 180)
    ctx.set_line_width(ls)
    ctx.set_source_rgb(grey, grey, grey)
    for a, b in lines:
        ctx.move_to(*pts[a]); ctx.line_to(*pts[b])
    ctx.stroke()
    ctx.restore()


def _hexagon(ctx, pts):                                     # draw.cpp:102-114
    ctx.set_source_rgb(1, 1, 1)
    ctx.move_to(*pts[0])
    for p in pts[1:6]:
        ctx.line_to(*p)
    ctx.close_path()
    ctx.stroke()


def _explosion(ctx, x, y, ls):                              # draw.cpp:116-145
    ctx.set_line_width(ls)
    ofs = 0
    for radius in range(15, 70, 8):
        ofs += 3
        g = .75 if radius < 60 else .5
        ctx.set_source_rgb(g, g, g)
        for angle in range(0, 360, 30):
            ctx.arc(x, y, radius, (angle + ofs) * math.pi / 180, (angle + ofs + 10) * math.pi / 180)
            ctx.stroke()
    ctx.set_source_rgb(.75, .75, .75)
    ctx.arc(x, y, 7, 0, math.pi * 2)
    ctx.stroke()


def draw_score(cairo, ctx, text, grey=.5):                  # draw.cpp:147-173
    ctx.select_font_face("monospace", cairo.FONT_SLANT_NORMAL, cairo.FONT_WEIGHT_BOLD)
    ctx.set_font_size(30)
    ctx.set_source_rgb(grey, grey, grey)
    ext = ctx.text_extents(text)
    w, h = (ext.width, ext.height) if hasattr(ext, "width") else (ext[2], ext[3])
    ctx.move_to(355 - w / 2.0, (290 - 193) + h / 2.0)
    ctx.show_text(text)


def new_context(cairo):
    surface = cairo.ImageSurface(cairo.FORMAT_RGB24, 90, 92)   # draw.cpp:62
    ctx = cairo.Context(surface)
    ctx.scale(90 / 450.0, 92 / 460.0)                           # draw.cpp:259-260
    ctx.translate(-130, -80)
    ctx.set_line_width(3)
    ctx.set_source_rgb(0, 0, 0)
    ctx.paint()
    return surface, ctx


def grey_of(surface):
    import numpy as np
    surface.flush()
    buf = np.frombuffer(surface.get_data(), np.uint8).reshape(92, surface.get_stride())[:, :360].reshape(92, 90, 4)
    return buf[..., 0].copy()                                   # B == G == R for grey input: RGBA2GRAY is the identity (ssf_env.py:205)


def render_state(cairo, s, ls=3):
    """s: a state record (oracle.Record / sf_state_record fields). Returns the (92, 90) uint8 grey frame."""
    surface, ctx = new_context(cairo)
    _hexagon(ctx, KA["hex_big"]); _hexagon(ctx, KA["hex_small"])            # draw.cpp:230-231
    if s.ship_alive:
        _wireframe(ctx, "ship", s.ship_x, s.ship_y, s.ship_angle, ls, 1)
    else:
        _explosion(ctx, s.ship_x, s.ship_y, ls)
    if s.fortress_alive:
        _wireframe(ctx, "fortress", 355, 315, s.fortress_angle, ls, 1)
    else:
        _explosion(ctx, 355, 315, ls)
    for i in range(20):
        if (s.missile_mask >> i) & 1:
            _wireframe(ctx, "missile", s.missile_x[i], s.missile_y[i], s.missile_angle[i], ls, 1)
    for i in range(20):
        if (s.shell_mask >> i) & 1 and math.hypot(s.shell_x[i] - 355, s.shell_y[i] - 315) > 21:
            _wireframe(ctx, "shell", s.shell_x[i], s.shell_y[i], s.shell_angle[i], ls, 1)
    draw_score(cairo, ctx, "%07d" % int(s.points))                          # draw.cpp:267
    kill = s.vulnerability > 10 and s.fortress_vuln_timer < 250               # draw.cpp:268
    ctx.set_line_width(ls - 1)                                              # draw.cpp:207-225
    ctx.set_source_rgb(.33, .33, .33); ctx.rectangle(255, 522, 200, 10); ctx.fill()
    g = 1 if kill else .66
    ctx.set_source_rgb(g, g, g); ctx.rectangle(255, 522, 20 * min(s.vulnerability, 10), 10); ctx.fill()
    return grey_of(surface)
