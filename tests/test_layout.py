"""CPU: the LAYOUT of the restated renderer (where the hexagons, the score and the vulnerability bar are) against the one
rendering the reference ships, rl/imgs/screens.png, through the committed measurements tests/golden/screens_layout.json
(tests/golden/make_layout_golden.py). The screenshots are full resolution (450 x 460) and resampled in the figure, so the
bar is a couple of user units = half a native pixel; anti-aliasing is not what this pins."""
import json
import os

import numpy as np

from oracle.oracle import OracleEnv, draw_native

LAYOUT = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "screens_layout.json")))


def user_box(mask):
    """bounding box of the lit native pixels in user units (viewport (130, 80, 450, 460) at scale 0.2: 5 units per pixel)"""
    ys, xs = np.nonzero(mask)
    return [130 + 5.0 * xs.min(), 80 + 5.0 * ys.min(), 130 + 5.0 * (xs.max() + 1), 80 + 5.0 * (ys.max() + 1)]


def test_layout_matches_the_reference_screenshots():
    r = OracleEnv("youturn", 1).get_state()
    r.ship_alive = 0; r.ship_x = -500.0; r.ship_y = -500.0   # nothing but the static layers
    r.vulnerability = 10; r.points = 8888888.0
    img = draw_native(r).astype(int)
    tol = 6.0   # user units: one native pixel (5) + the figure's resampling
    for p in LAYOUT["panels"]:
        big = user_box(img[:, :] > 100)
        # hexagons: white strokes; the big one bounds everything lit between rows 10 and 85
        hexes = np.zeros_like(img, bool); hexes[10:85] = img[10:85] > 60
        assert np.allclose(user_box(hexes), p["big_hex"], atol=tol), (user_box(hexes), p["big_hex"])
        small = np.zeros_like(img, bool); small[38:56, 35:55] = img[38:56, 35:55] > 60
        assert np.allclose(user_box(small), p["small_hex"], atol=tol), (user_box(small), p["small_hex"])
        bar = np.zeros_like(img, bool); bar[86:] = img[86:] > 40
        assert np.allclose(user_box(bar), p["bar"], atol=tol), (user_box(bar), p["bar"])
        # score: font dependent in shape, but its box (7 digits of a 30-unit monospace bold face centred at (355, 97)) is not
        score = np.zeros_like(img, bool); score[:9] = img[:9] > 30
        got, exp = user_box(score), p["score"]
        assert abs((got[0] + got[2]) / 2 - (exp[0] + exp[2]) / 2) <= tol and abs((got[1] + got[3]) / 2 - (exp[1] + exp[3]) / 2) <= tol, (got, exp)
        assert abs((got[2] - got[0]) - (exp[2] - exp[0])) <= 3 * tol and abs((got[3] - got[1]) - (exp[3] - exp[1])) <= 2 * tol, (got, exp)
