"""Generates tests/golden/screens_layout.json from the ONE rendering the reference ships: rl/imgs/screens.png (three full
resolution colour screenshots, 450 x 460 viewport pixels each, pasted side by side with margins). For each panel the
bounding boxes of the green hexagon strokes, of the vulnerability bar (grey / white) and of the score text (grey) are
measured and converted to the game's user units through the panel's own size (viewport (130, 80, 450, 460),
ssf_env.py:50). It pins the LAYOUT of the restated renderer (where things are), not its anti-aliasing.
usage: python tests/golden/make_layout_golden.py"""
import json, os
import cv2, numpy as np
img = cv2.imread("/root/reference/rl/imgs/screens.png", cv2.IMREAD_COLOR)   # BGR
H, W = img.shape[:2]
nonblack = (img.max(2) > 40)
cols = nonblack.any(0)
# panels = maximal runs of columns that contain the black game area: split at the white margins
white = (img.min(2) > 200).all(0)
edges, inside, start = [], False, 0
for x in range(W):
    if not white[x] and not inside:
        inside, start = True, x
    if (white[x] or x == W - 1) and inside:
        inside = False
        if x - start > 200:
            edges.append((start, x if white[x] else x + 1))
out = {"image": "rl/imgs/screens.png", "size": [W, H], "panels": []}
for (x0, x1) in edges:
    p = img[:, x0:x1]
    rows = np.nonzero(~(p.min(2) > 200).all(1))[0]
    p = p[rows.min():rows.max() + 1]
    ph, pw = p.shape[:2]
    sx, sy = 450.0 / pw, 460.0 / ph
    b, g, r = p[..., 0].astype(int), p[..., 1].astype(int), p[..., 2].astype(int)
    green = (g > 150) & (r < 100) & (b < 100)
    grey = (abs(r - g) < 12) & (abs(g - b) < 12) & (r > 60)
    grey[:4] = False; grey[-4:] = False; grey[:, :4] = False; grey[:, -4:] = False   # the figure's own frame lines
    def box(mask):
        ys, xs = np.nonzero(mask)
        return [130 + xs.min() * sx, 80 + ys.min() * sy, 130 + (xs.max() + 1) * sx, 80 + (ys.max() + 1) * sy]
    gy, gx = np.nonzero(green)
    big = box(green)
    # small hexagon = green pixels within 60 user units of the centre (355, 315)
    ux, uy = 130 + gx * sx, 80 + gy * sy
    near = (abs(ux - 355) < 60) & (abs(uy - 315) < 60)
    small = [ux[near].min(), uy[near].min(), ux[near].max() + sx, uy[near].max() + sy]
    top = grey.copy(); top[int(60 / sy):] = False       # score text: grey pixels in the top 60 user units
    bot = grey.copy(); bot[:int(420 / sy)] = False      # bar: grey pixels in the bottom 40
    out["panels"].append({"panel_px": [int(x0), int(rows.min()), int(pw), int(ph)], "big_hex": big, "small_hex": small, "score": box(top), "bar": box(bot)})
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "screens_layout.json"), "w"), indent=1)
for p in out["panels"]:
    print({k: [round(v, 1) for v in p[k]] for k in ("big_hex", "small_hex", "score", "bar")})
