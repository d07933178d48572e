"""Host-side rendering of env 0 of the car-flag env.

The reference draws a 48 x 600 number line (gym_po/envs/car_flag.py:146-185; its pyglet branch is disabled,
``visualize = None`` :16-19): white end posts, heaven flag green, hell flag red, the priest zone as three blue
bars, the car as a grey block on the lower half (white once the priest indicator is set).  Only env 0's three
state floats and two flag bits are copied from the device.  Pixel-exact vs frames drawn by the real reference
(tests/golden/render_car.npz).
"""
from __future__ import annotations

import numpy as np

WIDTH, BAR, CAR_H = 600, 4, 24          # SCREEN_WIDTH, PIXEL_WIDTH, PIXEL_HEIGHT (car_flag.py:37-43)
MAX_POS, PRIEST_ZONE = 1.1, 0.2


def to_pixel(x):
    """Start column of a 4-pixel bar at position x: floor of the linear map [-1.1, 1.1] -> [0, 596] (:82-87)."""
    return np.floor(np.interp(x, [-MAX_POS, MAX_POS], [0, WIDTH - BAR])).astype(int)


def render_car(pos: float, indicator: float, heaven: float, priest: float) -> np.ndarray:
    img = np.zeros((2 * CAR_H, WIDTH, 3), dtype=np.uint8)
    img[:, :BAR] = 255
    img[:, -BAR:] = 255
    left, right = to_pixel([-1, 1])
    good, bad = (left, right) if heaven < 0 else (right, left)
    img[:, good:good + BAR, 1] = 255
    img[:, bad:bad + BAR, 0] = 255
    car = int(to_pixel(pos))
    img[-CAR_H:, car:car + BAR] = 255 if indicator else 128
    lo, mid, hi = to_pixel([priest - PRIEST_ZONE, priest, priest + PRIEST_ZONE])
    img[:, lo:lo + BAR, 2] = 128
    img[:, hi:hi + BAR, 2] = 128
    img[:, mid:mid + BAR, 2] = 255
    return img
