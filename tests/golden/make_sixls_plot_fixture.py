#!/usr/bin/env python
"""Digitise the ONE artefact of the real ACE binary the reference holds for this path: ``pyaceqd/tests/sixls_compare.png``,
the plot its author kept to compare new runs of ``pyaceqd/tests/six_level_linear.py`` against (the script saves
``sixls_compare_.png`` next to it).  Six-level system, ARP + TPE pulses, LA phonons at 4 K, in-plane field 2 T, dt = 0.1 ps,
-60 ... 180 ps: the occupations g, x, y, s, f, b as matplotlib default-colour lines on default axes.

The figure is matplotlib's default 640x480 canvas (axes box [0.125, 0.11, 0.775, 0.77] -> pixel columns 80 ... 576, rows
58 ... 427: checked below against the black frame), the data limits are the plotted ranges plus 5 % margins
(t in [-60, 180] -> [-72, 192]; occupations in [0, 1] -> [-0.05, 1.05]).  For every pixel column and curve colour the
script stores the data-coordinate range of the rows that carry that colour.  Resolution: 0.53 ps and 0.003 per pixel.

Run in the build container (needs /root/reference and PIL); writes tests/golden/reference_sixls_plot.json."""
import json
import os
import sys

import numpy as np
from PIL import Image

SRC = "/root/reference/pyaceqd/tests/sixls_compare.png"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_sixls_plot.json")
COLORS = {"g": (0x1f, 0x77, 0xb4), "x": (0xff, 0x7f, 0x0e), "y": (0x2c, 0xa0, 0x2c), "s": (0xd6, 0x27, 0x28),
          "f": (0x94, 0x67, 0xbd), "b": (0x8c, 0x56, 0x4b)}          # C0 ... C5 in plotting order
W, H = 640, 480
AX = (0.125 * W, 0.11 * H, 0.775 * W, 0.77 * H)                       # left, bottom (from below), width, height
XLIM, YLIM = (-72.0, 192.0), (-0.05, 1.05)


def main():
    im = np.array(Image.open(SRC).convert("RGB")).astype(int)
    assert im.shape == (H, W, 3), im.shape
    dark = im.sum(axis=2) < 60
    cols = [i for i, c in enumerate(dark.sum(axis=0)) if c > 300]
    rows = [i for i, c in enumerate(dark.sum(axis=1)) if c > 300]
    assert cols == [80, 576] and rows == [58, 427], (cols, rows)     # the axes frame sits where the defaults put it
    t_of = lambda px: XLIM[0] + (px + 0.5 - AX[0]) / AX[2] * (XLIM[1] - XLIM[0])
    v_of = lambda row: YLIM[0] + (H - (row + 0.5) - AX[1]) / AX[3] * (YLIM[1] - YLIM[0])
    curves = {}
    for name, rgb in COLORS.items():
        near = np.abs(im - np.array(rgb)).sum(axis=2) < 24          # core pixels of the line, not its anti-aliased rim
        pts = []
        for px in range(81, 576):
            rr = np.nonzero(near[:, px])[0]
            rr = rr[(rr > 58) & (rr < 427)]
            if px >= 500 and len(rr):                                 # the legend box (upper right) repeats the colours
                rr = rr[rr > 200]
            if len(rr) == 0:
                continue
            pts.append([round(t_of(px), 4), round(v_of(rr.max()), 5), round(v_of(rr.min()), 5)])   # t, v_low, v_high
        curves[name] = pts
        print(name, len(pts), "columns", file=sys.stderr)
    json.dump({"source": "pyaceqd/tests/sixls_compare.png (plot made by the reference with the real ACE binary)",
               "script": "pyaceqd/tests/six_level_linear.py:5-10", "pixel_dt": (XLIM[1] - XLIM[0]) / AX[2],
               "pixel_dv": (YLIM[1] - YLIM[0]) / AX[3], "curves": curves}, open(OUT, "w"))
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
