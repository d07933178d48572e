"""ROOMS layouts: text asset -> integer grid (host side, parsed once at construction).

Same public names as the reference module (rooms/layouts.py: ``LAYOUTS``, ``ENDS``, ``STARTS``,
``layout_to_np``, ``np_to_grid``); the maps live in ``layouts.txt`` next to this file.
"""
from __future__ import annotations

import os

import numpy as np

__all__ = ["LAYOUTS", "layout_to_np", "np_to_grid", "ENDS", "STARTS", "WALL_CHAR"]

WALL_CHAR = "x"
LAYOUTS, ENDS, STARTS = {}, {}, {}


def _load():
    name = None
    rows = {}
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "layouts.txt")) as f:
        for raw in f:
            line = raw.strip()
            if not line or line[0] == "#":
                continue
            if line[0] == "@":
                head = line[1:].split()
                name = head[0]
                rows[name] = []
                for item in head[1:]:
                    key, val = item.split("=")
                    x, y = (int(v) for v in val.split(","))
                    {"end": ENDS, "start": STARTS}[key][name] = (x, y)
            else:
                rows[name].append(line)
    for k, v in rows.items():
        LAYOUTS[k] = "\n".join(v)


_load()


def layout_to_np(layout: str) -> np.ndarray:
    """Layout string -> 2-D array of single characters."""
    return np.array([list(r.strip()) for r in layout.splitlines()])


def np_to_grid(chars: np.ndarray) -> np.ndarray:
    """Characters -> int grid: -1 for walls, otherwise the rank of the cell's room character among the
    sorted distinct room characters."""
    rooms = sorted(set(chars.ravel().tolist()) - {WALL_CHAR})
    lut = {ch: i for i, ch in enumerate(rooms)}
    lut[WALL_CHAR] = -1
    return np.vectorize(lut.__getitem__, otypes=[np.int64])(chars)
