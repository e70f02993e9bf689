#!/usr/bin/env python
"""Fixtures for the reference's end-to-end JPEG regression number:
`RMSError,1.71767` in /root/reference/benchmark/results.txt:37-42 -- compare_jpeg_jojpeg
(homo/fhe_image.h:508-521) between the picture decoded from the homomorphically computed JPEG and the
picture decoded from jo_write_jpg(original, QUALITY = 0), for image/boazbarak.jpg (48 x 48).

Run in the container that has /root/reference:  python tests/golden/make_golden_jpeg.py
It uses oracle/_ref/ref_jpeg_harness (`make -C oracle ref`: the reference's own stb_image decoder and
jo_jpeg encoder, compiled from /root/reference/include) and writes
  tests/golden/boazbarak_rgb.bin        48*48*3 bytes: stbi_load(image/boazbarak.jpg)
  tests/golden/boazbarak_jo_q0_rgb.bin  48*48*3 bytes: stbi_load(jo_write_jpg(..., quality 0))
  tests/golden/jpeg_golden.json         sizes, the published number and where it is published
"""
import json
import os
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
H = os.path.join(ROOT, "oracle", "_ref", "ref_jpeg_harness")
REF = "/root/reference"

subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"])
with tempfile.TemporaryDirectory() as tmp:
    orig, jo_jpg, jo_rgb = (os.path.join(tmp, f) for f in ("orig.rgb", "jo.jpg", "jo.rgb"))
    w, h = map(int, subprocess.check_output([H, "decode", os.path.join(REF, "image", "boazbarak.jpg"), orig]).split())
    subprocess.check_call([H, "jo", orig, str(w), str(h), "0", jo_jpg])          # QUALITY = 0, fhe_image.h:32
    subprocess.check_call([H, "decode", jo_jpg, jo_rgb])
    for src, name in ((orig, "boazbarak_rgb.bin"), (jo_rgb, "boazbarak_jo_q0_rgb.bin")):
        with open(src, "rb") as f, open(os.path.join(HERE, name), "wb") as g:
            g.write(f.read())
# resize comparison pictures (homo/fhe_resize.h:35-68 compare_resize_opencv): cv::resize of
# cv::imread(original) to 17 x 17 (benchmark/benchmark.py:9), stored RGB
import cv2
import numpy as np
bgr = cv2.imread(os.path.join(REF, "image", "boazbarak.jpg"), cv2.IMREAD_COLOR)
for flag, name in ((cv2.INTER_LINEAR, "boazbarak_cv_linear_17_rgb.bin"), (cv2.INTER_CUBIC, "boazbarak_cv_cubic_17_rgb.bin")):
    np.ascontiguousarray(cv2.resize(bgr, (17, 17), interpolation=flag)[..., ::-1]).tofile(os.path.join(HERE, name))
json.dump({"image": "image/boazbarak.jpg", "width": w, "height": h,
           "resize": {"width": 17, "height": 17, "published_rms_error_bilinear": "17.9597",
                      "published_rms_error_bicubic": "19.8048",
                      "published_at": "benchmark/results.txt (30 bilinear rows, e.g. :17-21; 20 bicubic rows, e.g. :22-26)",
                      "comparator": "homo/fhe_resize.h:35-68 against cv::resize(cv::imread(original)); OpenCV " + cv2.__version__ + " here"},
           "published_rms_error": "1.71767", "published_at": "benchmark/results.txt:37-42 (every t >= 3001, every n)",
           "comparator": "homo/fhe_image.h:508-521 (int accumulator of squared uint8 differences, sqrt of the mean, printed with operator<<)",
           "decoder": "stb_image (reference include/stb_image.h)", "encoder_of_the_comparison_picture": "jo_write_jpg quality 0 -> 90 (include/jo_jpeg.h:251)"},
          open(os.path.join(HERE, "jpeg_golden.json"), "w"), indent=1)
print("wrote fixtures for", w, "x", h)
