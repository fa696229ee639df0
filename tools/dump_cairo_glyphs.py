"""Render the score digits with REAL cairo and the machine's "monospace" bold face (what drawScore, draw.cpp:160-173, does)
and write them in the format sf_set_glyph_masks (include/sf_b200.h) / SFVecEnv.set_glyph_masks takes: alpha uint8
[10, 5, 27] = coverage of digit d drawn in all 7 slots over native rows 1..5 x columns 32..58, slot uint8 [27] = the slot
owning each strip column. Needs pycairo or cairocffi and a font (neither is in the build image: the script then says so
and exits 0). usage: python tools/dump_cairo_glyphs.py out.npz"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np
import cairo_ref

X0, Y0, W, H = 32, 1, 27, 5   # SF_TEXT_* of csrc/sf_tables.h


def strip_of(cairo, text):
    surface, ctx = cairo_ref.new_context(cairo)
    cairo_ref.draw_score(cairo, ctx, text, grey=1.0)   # white on black: the pixel value IS the glyph coverage
    img = cairo_ref.grey_of(surface)
    outside = img.copy(); outside[Y0:Y0 + H, X0:X0 + W] = 0
    if outside.any():
        ys, xs = np.nonzero(outside)
        raise SystemExit("this font's digits leave the strip the kernels are built for (rows %d..%d, columns %d..%d lit): "
                         "SF_TEXT_X0/Y0/W/H in csrc/sf_tables.h must grow first" % (ys.min(), ys.max(), xs.min(), xs.max()))
    return img[Y0:Y0 + H, X0:X0 + W]


def main():
    cairo = cairo_ref.import_cairo()
    if cairo is None:
        print("dump_cairo_glyphs: neither pycairo nor cairocffi can be imported here -> nothing written")
        return 0
    alpha = np.stack([strip_of(cairo, str(d) * 7) for d in range(10)])
    slot = np.full(W, 255, np.uint8)
    for k in range(7):   # which columns does slot k light (widest digit, spaces elsewhere: a monospace face advances the same)
        cols = np.nonzero(strip_of(cairo, " " * k + "8" + " " * (6 - k)).any(0))[0]
        for d in "0123456789":
            cols = np.union1d(cols, np.nonzero(strip_of(cairo, " " * k + d + " " * (6 - k)).any(0))[0])
        if (slot[cols] != 255).any():
            raise SystemExit("two digits share a pixel column at this size: the per-column slot table cannot describe this font")
        slot[cols] = k
    np.savez_compressed(sys.argv[1] if len(sys.argv) > 1 else "glyph_masks.npz", alpha=alpha, slot=slot)
    print("wrote masks of shape", alpha.shape, "slots", slot.tolist())
    return 0


if __name__ == "__main__":
    sys.exit(main())
