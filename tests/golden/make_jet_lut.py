"""Writes tests/golden/jet_lut_rgb.npy: cv2.COLORMAP_JET as a [256, 3] RGB table (what the reference's create_heatmap,
evaluate_video.py:60-61, applies).  The same table is compiled into vad_api.cu (c_jet_rgb)."""
import os

import cv2
import numpy as np

lut = cv2.applyColorMap(np.arange(256, dtype=np.uint8).reshape(1, 256), cv2.COLORMAP_JET)
rgb = cv2.cvtColor(lut, cv2.COLOR_BGR2RGB).reshape(256, 3).astype(np.uint8)
np.save(os.path.join(os.path.dirname(os.path.abspath(__file__)), "jet_lut_rgb.npy"), rgb)
