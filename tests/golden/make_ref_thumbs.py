"""Signatures of the renders the reference ships (SURVEY 8c iii) — the only outputs of the real reference that exist.
Run in the build container (it reads /root/reference, which the GPU box does not have):

    python tests/golden/make_ref_thumbs.py

Writes tests/golden/ref_thumbs.npz:
  lights_rows, banner_rows  (rows, 9, 3) int8, -1 padded — light-panel colour codes, far row to near row, left to right
                            (tests/panel_rows.py) of public_html/images/banners/lights.png and banner.png (museum scene)
  bunny                     (18, 30, 3) float32 — 10x10-pixel block means of public_html/images/banners/bunny_high.png
"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from panel_rows import panel_rows  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(HERE, "ref_thumbs.npz")


def load(path):
    return cv2.imread(path, cv2.IMREAD_UNCHANGED)[..., :3][..., ::-1].astype(np.float32) / 255.0   # BGR(A) -> RGB


def grid(im, gh=18, gw=30):
    h, w = im.shape[:2]
    return im[: h // gh * gh, : w // gw * gw].reshape(gh, h // gh, gw, w // gw, 3).mean((1, 3)).astype(np.float32)


def pack(rows):
    a = -np.ones((len(rows), 9, 3), np.int8)
    for i, r in enumerate(rows):
        a[i, : len(r)] = np.array(r, np.int8)
    return a


if __name__ == "__main__":
    lights = panel_rows(load(os.path.join(REF, "public_html/images/banners/lights.png")))
    banner = panel_rows(load(os.path.join(REF, "banner.png")))
    np.savez(OUT, lights_rows=pack(lights), banner_rows=pack(banner), bunny=grid(load(os.path.join(REF, "public_html/images/banners/bunny_high.png"))))
    for name, rows in (("lights.png", lights), ("banner.png", banner)):
        print(name)
        for r in rows:
            print("  ", len(r), r)
