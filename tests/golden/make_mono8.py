#!/usr/bin/env python
"""Generates tests/golden/mono8_cv4.npz: what this image's OpenCV (cv2 4.x) returns for cv2.cvtColor(..., COLOR_*2GRAY)
on small 8-bit images -- the pin for the fixed-point arithmetic of oracle `orc_to_mono8` (rule 1 = OpenCV 4 constants;
rule 0 differs only in its three published constants and the shift).  The first image walks every (B, G, R) on a coarse
lattice plus the extremes, the second is random.
    python tests/golden/make_mono8.py"""
import os

import cv2
import numpy as np

rng = np.random.RandomState(77)
lattice = np.array([[b, g, r] for b in range(0, 256, 15) for g in range(0, 256, 15) for r in range(0, 256, 15)], np.uint8)
pad = (-len(lattice)) % 64
lattice = np.concatenate([lattice, rng.randint(0, 256, (pad, 3)).astype(np.uint8)]).reshape(-1, 64, 3)
rand = rng.randint(0, 256, (37, 53, 3)).astype(np.uint8)
out = {"cv2_version": np.array(cv2.__version__)}
for name, img in (("lattice", lattice), ("rand", rand)):
    alpha = rng.randint(0, 256, img.shape[:2] + (1,)).astype(np.uint8)
    out[name + "_bgr8"] = img
    out[name + "_bgra8"] = np.concatenate([img, alpha], axis=2)
    out[name + "_gray_from_bgr8"] = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
    out[name + "_gray_from_rgb8"] = cv2.cvtColor(img, cv2.COLOR_RGB2GRAY)
    out[name + "_gray_from_bgra8"] = cv2.cvtColor(out[name + "_bgra8"], cv2.COLOR_BGRA2GRAY)
    out[name + "_gray_from_rgba8"] = cv2.cvtColor(out[name + "_bgra8"], cv2.COLOR_RGBA2GRAY)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "mono8_cv4.npz"), **out)
print("wrote mono8_cv4.npz with cv2", cv2.__version__)
