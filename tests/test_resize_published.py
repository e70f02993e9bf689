"""The end-to-end numbers the reference publishes for the resize path: RMSError of the
homomorphically resized image/boazbarak.jpg (48 x 48 -> 17 x 17) against OpenCV's resize
(homo/fhe_resize.h:35-68), benchmark/results.txt: 17.9597 bilinear (30 parameter sets), 19.8048 bicubic.

The plaintext mirror of ResizeImage / SampleLinear / Linear at HEAD (float coordinates, truncation
toward zero, clamped neighbourhoods, int(decode) and CLAMP on the client) reproduces the bilinear number
digit for digit.  The bicubic number is NOT reproduced by HEAD's graph -- the mirror of
Cubic as written (incl. t3 = t * t, fhe_resize.h:175) gives 21.0333 and no single-edit variant of the
sampler gives 19.8048 -- so the published sweep predates some change to Cubic; it is recorded here as
a known difference, not asserted.  The GPU version of the bilinear check runs the whole encrypted
pipeline (tests/test_server_resize_decode.py)."""
import json
import os

import numpy as np

from conftest import ROOT

GOLDEN = os.path.join(ROOT, "tests", "golden")
f32 = np.float32


def fixtures():
    meta = json.load(open(os.path.join(GOLDEN, "jpeg_golden.json")))
    img = np.fromfile(os.path.join(GOLDEN, "boazbarak_rgb.bin"), dtype=np.uint8).reshape(meta["height"], meta["width"], 3)
    r = meta["resize"]
    shape = (r["height"], r["width"], 3)
    lin = np.fromfile(os.path.join(GOLDEN, "boazbarak_cv_linear_17_rgb.bin"), dtype=np.uint8).reshape(shape)
    cub = np.fromfile(os.path.join(GOLDEN, "boazbarak_cv_cubic_17_rgb.bin"), dtype=np.uint8).reshape(shape)
    return img, lin, cub, r


def sample_coordinates(ow, oh, dw, dh):
    """u, v of every output pixel as ResizeImage computes them (fhe_resize.h:351, 382)"""
    vs = [f32(np.float64(f32(y) / f32(dh - 1) * f32(oh)) - 0.5) for y in range(dh)]
    us = [f32(np.float64(f32(x) / f32(dw - 1) * f32(ow)) - 0.5) for x in range(dw)]
    return us, vs


def client_pixels(values):
    """client_resize --recieve (client_resize.cpp:201-210): int pixel = decode; CLAMP(pixel, 0, 255)"""
    return np.clip(np.trunc(np.asarray(values, dtype=np.float64)), 0, 255).astype(np.uint8)


def reference_rms(a, b):
    d = a.astype(np.int64) - b.astype(np.int64)
    return format(float(np.sqrt(float((d * d).sum()) / d.size)), "g")


def resize_mirror(img, dw, dh, bicubic):
    oh, ow, _ = img.shape
    src = img.astype(np.float64)

    def px(x, y):
        return src[min(max(y, 0), oh - 1), min(max(x, 0), ow - 1)]

    def linear(A, B, t):
        return (1.0 - t) * A + t * B

    def cubic(A, B, C, D, t):
        a, b, c = -A + 3 * B - 3 * C + D, 2 * A - 5 * B + 4 * C - D, C - A
        return (a * (t * t) + b * (t * t) + c * t) * 0.5 + B          # t3 = t * t (sic)

    us, vs = sample_coordinates(ow, oh, dw, dh)
    out = np.zeros((dh, dw, 3))
    for y, v in enumerate(vs):
        yi, yf = int(v), float(v - np.floor(v))
        for x, u in enumerate(us):
            xi, xf = int(u), float(u - np.floor(u))
            if bicubic:
                cols = [cubic(px(xi - 1, yi - 1 + r), px(xi, yi - 1 + r), px(xi + 1, yi - 1 + r), px(xi + 2, yi - 1 + r), xf) for r in range(4)]
                out[y, x] = cubic(*cols, yf)
            else:
                out[y, x] = linear(linear(px(xi, yi), px(xi + 1, yi), xf), linear(px(xi, yi + 1), px(xi + 1, yi + 1), xf), yf)
    return out


def test_bilinear_published_rms_error():
    img, lin, _, r = fixtures()
    got = client_pixels(resize_mirror(img, r["width"], r["height"], bicubic=False))
    assert reference_rms(got, lin) == r["published_rms_error_bilinear"]          # 17.9597


def test_bicubic_head_graph_differs_from_the_published_sweep():
    img, _, cub, r = fixtures()
    got = client_pixels(resize_mirror(img, r["width"], r["height"], bicubic=True))
    rms = reference_rms(got, cub)
    assert rms == "21.0333" and rms != r["published_rms_error_bicubic"]
