"""CPU: frame oracle checks. The INTER_AREA restatement is pinned against real cv2; the product's
host-built static tables (background hexagons, fortress sprites, fortress explosion, digits, bar) are
compared bit-exactly with the oracle's independent rasteriser."""
import ctypes as C

import numpy as np
import pytest

from oracle.oracle import OracleEnv, Record, draw_native, draw_obs, resize_area
from spacefortress_b200 import _lib


def test_inter_area_restatement_matches_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.RandomState(0)
    for i in range(200):
        img = rng.randint(0, 256, (92, 90)).astype(np.uint8)
        if i % 3 == 0:
            img = ((rng.rand(92, 90) < 0.1) * rng.randint(0, 256, (92, 90))).astype(np.uint8)
        assert np.array_equal(cv2.resize(img, (84, 84), interpolation=cv2.INTER_AREA), resize_area(img)), i


def test_gray_conversion_is_identity_on_grey_pixels():
    cv2 = pytest.importorskip("cv2")
    v = np.arange(256, dtype=np.uint8)
    bgrx = np.stack([v, v, v, np.zeros_like(v)], axis=1).reshape(16, 16, 4)
    assert np.array_equal(cv2.cvtColor(bgrx, cv2.COLOR_RGBA2GRAY).reshape(-1), v)


def test_obs_is_resize_of_native():
    cv2 = pytest.importorskip("cv2")
    o = OracleEnv("youturn", 3)
    rng = np.random.RandomState(1)
    for t in range(150):
        o.step(int(rng.randint(16)))
        if t % 10 == 0:
            assert np.array_equal(cv2.resize(o.native_frame(), (84, 84), interpolation=cv2.INTER_AREA), o.obs())


def test_frame_layout_sanity():
    """Layout anchors from SURVEY.md R2/R8: hexagon extents and the vulnerability bar rows/cols."""
    r = Record(); r.ship_alive = 1; r.ship_x = -5000; r.ship_y = -5000; r.fortress_alive = 1; r.fortress_angle = 180
    f = draw_native(r)
    ys, xs = np.nonzero(f[8:86])  # between score strip and bar: hexagons + fortress
    assert 4 <= xs.min() <= 5 and 84 <= xs.max() <= 85
    assert (f[88:91, 25:65] > 0).all() and (f[88:91, :25] == 0).all() and (f[88:91, 65:] == 0).all()
    assert f[89, 30] == 84  # bar background .33
    r.vulnerability = 5
    f = draw_native(r)
    assert f[89, 30] == 168 and f[89, 25 + 20] == 84  # 5 * 4 px filled with .66
    r.vulnerability = 12; r.fortress_vuln_timer = 10
    assert draw_native(r)[89, 60] == 255


@pytest.mark.parametrize("alive", [1, 0])
def test_static_tables_match_oracle(alive):
    L = _lib.lib()
    nat = np.zeros((92, 90), np.uint8)
    bgo = np.zeros((84, 84), np.uint8)
    cases = ((0, 0, 0), (1234567, 7, 0), (89, 12, 1), (905, 10, 0), (9999999, 3, 0))
    for ang in range(0, 360, 10):
        for pts, v, kill in cases:
            assert L.sf_host_static_frame(alive, ang, pts, v, kill, nat.ctypes.data_as(C.c_void_p), bgo.ctypes.data_as(C.c_void_p)) == 0
            r = Record(); r.ship_alive = 1; r.ship_x = -5000; r.ship_y = -5000
            r.fortress_alive = alive; r.fortress_angle = ang; r.points = pts; r.vulnerability = v
            r.fortress_vuln_timer = 0 if kill else 300
            assert np.array_equal(draw_native(r), nat), (alive, ang, pts, v, kill)


def test_background_observation_table():
    """bg_obs (the 84x84 frame the kernel starts every observation from) == INTER_AREA(hexagons + score
    "0000000" + empty vulnerability bar), through the oracle's cv2-pinned resize."""
    L = _lib.lib()
    nat = np.zeros((92, 90), np.uint8)
    bgo = np.zeros((84, 84), np.uint8)
    assert L.sf_host_static_frame(-1, 0, 0, 0, 0, nat.ctypes.data_as(C.c_void_p), bgo.ctypes.data_as(C.c_void_p)) == 0
    assert nat.max() > 100 and (nat > 0).sum() > 300  # two hexagons (0.6 px lines never cover a whole pixel)
    assert np.array_equal(resize_area(nat), bgo)
