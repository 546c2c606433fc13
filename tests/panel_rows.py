"""Light-panel colour sequences of a museum render — the coarse, camera-independent signature used to compare our
renders with the images the reference ships (tests/golden/make_ref_thumbs.py, tests/test_reference_images.py).

The museum's 3 x 9 area lights (src/scenes.rs:30-40,54-68) are directly visible emitters of intensity colour x 2.5, so a
render shows each as a saturated blob whose clamped colour is one of nine codes: per channel 2 = full (>= 0.9),
1 = partial (the 0.3 x 2.5 = 0.75 of the pastel colours), 0 = off. Blobs are grouped into rows (far to near) and read
left to right; adjacent duplicates (a blob split by noise) are merged."""
import cv2
import numpy as np


def panel_rows(rgb, min_area=None):
    """rgb: float array (h, w, 3) in 0..1 -> list of rows (top to bottom), each a list of (r, g, b) codes left to right."""
    if min_area is None:
        min_area = max(12, int(0.00012 * rgb.shape[0] * rgb.shape[1]))
    mask = (rgb.max(-1) > 0.93).astype(np.uint8)
    n, lab, stats, cent = cv2.connectedComponentsWithStats(mask, connectivity=8)
    blobs = []
    for i in range(1, n):
        x, y, w, h, area = stats[i]
        if area < min_area or w < 4 or area < 0.42 * w * h:      # light panels are solid quads; firefly clusters are not
            continue
        c = rgb[lab == i].mean(0)
        blobs.append((float(cent[i][1]), float(cent[i][0]), tuple(2 if v > 0.9 else (1 if v > 0.45 else 0) for v in c)))
    blobs.sort()
    rows = []
    for b in blobs:
        if rows and abs(rows[-1][-1][0] - b[0]) < max(6.0, 0.012 * rgb.shape[0]):
            rows[-1].append(b)
        else:
            rows.append([b])
    out = []
    for r in rows:
        seq = []
        for _, _, code in sorted(r, key=lambda t: t[1]):
            if not seq or seq[-1] != code:
                seq.append(code)
        if len(seq) >= 5:                      # partial rows at the image border carry no order information
            out.append(seq)
    return out


def contains(row, sub):
    """True if `sub` is a contiguous run of `row`."""
    return any(row[i:i + len(sub)] == sub for i in range(len(row) - len(sub) + 1))
