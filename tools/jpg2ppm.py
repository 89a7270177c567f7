#!/usr/bin/env python
"""Converts the images of a directory to binary PPM (P6) next to the originals, for examples/eval_ffd.cpp (this image has
no JPEG decoder for C++; the reference uses cv::imread).  usage: jpg2ppm.py DIR [OUT_DIR]"""
import sys
from pathlib import Path

import cv2

src = Path(sys.argv[1]); dst = Path(sys.argv[2]) if len(sys.argv) > 2 else src
dst.mkdir(parents=True, exist_ok=True)
n = 0
for p in sorted(src.iterdir()):
    if p.suffix.lower() in (".jpg", ".jpeg", ".png", ".bmp"):
        img = cv2.imread(str(p), cv2.IMREAD_COLOR)
        if img is not None:
            cv2.imwrite(str(dst / (p.stem + ".ppm")), img); n += 1
print(n, "images converted to", dst)
