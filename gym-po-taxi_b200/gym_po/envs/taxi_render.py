"""Host-side rendering of Taxi device state.

``TaxiVecEnv.render(idx)`` of the reference (gym_po/envs/extended_taxi.py:289-342 with ``str_map_to_img``
:121-146 and ``tile_images``, gym_po/envs/render_utils.py:63-88) draws the selected envs as coloured cell
grids, tiles them, rescales with ``cv2.INTER_AREA`` and prints env 0's last action in a 20-pixel strip.  Here
only the encoded states of the selected envs are copied from the device; the frame is produced on the host from a
palette-index grid.  Pixel output equals the reference's for ``idx = arange(k)`` (tests/golden/render_taxi_*.npz;
for other ``idx`` the reference indexes its frame stack with env ids and fails or mis-draws — here each selected
env gets its own tile).

Reference behaviour kept: a passenger waiting on the taxi's cell shows as the taxi ("TP" is stored into a one-char
array, i.e. "T"); the Hansen highlight adds 64 (mod 256) to the four cells around the taxi; the frame is resized
to (map_cols*16 rows) x (map_rows*16 columns) whatever the tiling.
"""
from __future__ import annotations

import numpy as np

CELL_PX = 16
TEXT_STRIP = 20
# palette rows: wall, floor, pseudo-wall ':', named location, destination, taxi, passenger, full taxi
PALETTE = np.array([[0, 0, 0], [96, 96, 96], [0, 128, 128], [191, 191, 191], [0, 0, 128], [128, 128, 0], [128, 0, 128],
                    [0, 128, 0]], dtype=np.uint8)
K_WALL, K_FLOOR, K_PSEUDO, K_LOC, K_DEST, K_TAXI, K_PASS, K_FULL = range(8)


def map_codes(rows):
    """'|'-bordered palette-index grid of the map strings and the cell -> grid-coordinate scale (1 or 2)."""
    sep = any(":" in r for r in rows)
    body = ["|" + r + "|" for r in rows]
    full = ["|" * len(body[0])] + body + ["|" * len(body[0])]
    lut = {"|": K_WALL, " ": K_FLOOR, ":": K_PSEUDO}
    codes = np.array([[lut.get(ch, K_LOC) for ch in line] for line in full], dtype=np.uint8)
    return codes, (2 if sep else 1)


def taxi_tiles(codes, scale, r, c, p, d, np_locs, nlocs, hansen):
    """[k, H, W, 3] uint8 tiles for k envs given decoded state components (host arrays)."""
    k = len(r)
    gy = lambda y: np.asarray(y) + 1
    gx = (lambda x: 2 * np.asarray(x) + 1) if scale == 2 else (lambda x: np.asarray(x) + 1)
    grid = np.repeat(codes[None], k, 0)
    n = np.arange(k)
    ty, tx = gy(r), gx(c)
    grid[n, gy(np_locs[d, 0]), gx(np_locs[d, 1])] = K_DEST
    grid[n, ty, tx] = K_TAXI
    waiting = p < nlocs
    py, px = gy(np_locs[p, 0]), gx(np_locs[p, 1])
    on_taxi = waiting & (py == ty) & (px == tx)
    show = waiting & ~on_taxi
    grid[n[show], py[show], px[show]] = K_PASS
    grid[n[~waiting], ty[~waiting], tx[~waiting]] = K_FULL
    img = PALETTE[grid]
    if hansen:   # the four grid neighbours of the taxi get +64 (uint8 wrap-around), once each
        for dy, dx in ((-1, 0), (1, 0), (0, -1), (0, 1)):
            img[n, ty + dy, tx + dx] += 64
    return img


def tile(img_nhwc):
    """N tiles -> one P x Q sheet, P = ceil(sqrt(N)), Q = ceil(N / P), row-major, black padding."""
    n, h, w, ch = img_nhwc.shape
    p = int(np.ceil(np.sqrt(n)))
    q = int(np.ceil(n / p))
    sheet = np.zeros((p * q, h, w, ch), dtype=img_nhwc.dtype)
    sheet[:n] = img_nhwc
    return sheet.reshape(p, q, h, w, ch).swapaxes(1, 2).reshape(p * h, q * w, ch)


def render_taxi(map_rows, s, nlocs, np_locs, cols, hansen, last_action_name=None):
    """RGB frame for the encoded states ``s`` (host int array) — see the module docstring."""
    import cv2
    codes, scale = map_codes(map_rows)
    s = np.asarray(s, dtype=np.int64)
    d = s % nlocs
    t = s // nlocs
    p = t % (nlocs + 1)
    t = t // (nlocs + 1)
    tiles = taxi_tiles(codes, scale, t // cols, t % cols, p, d, np.asarray(np_locs), nlocs, hansen)
    h, w = codes.shape
    frame = cv2.resize(tile(tiles), (h * CELL_PX, w * CELL_PX), interpolation=cv2.INTER_AREA)   # dsize = (width, height)
    frame = np.concatenate((frame, np.zeros((frame.shape[0], TEXT_STRIP, 3), dtype=np.uint8)), axis=1)
    if last_action_name is not None:
        cv2.putText(frame, f"  ({last_action_name})\n", (5, frame.shape[1] - TEXT_STRIP), cv2.FONT_HERSHEY_SIMPLEX, 0.25,
                    (255, 255, 255), 1, lineType=cv2.LINE_AA)
    return frame
