"""Derives a golden hit/miss mask from the reference's own checked-in render /root/reference/teapot_4k_tris.png.

The PNG predates the current reference code (older sky constant, Solid teapot, thinner wire-frame edges and
a different upper-left disk), so it is not a pixel golden.  What it still pins is the GEOMETRY of the path as the
reference's own binary drew it: the camera conventions (row -> -x, column -> +y, 90 degree fov, eye behind the viewport
plane), parse_obj's transform of the teapot, and the lower-right mirror disk with its reflection of the teapot.
Every 8th pixel (centre of each 8x8 block) is classified sky / not sky and stored as packed bits.

    python tests/golden/make_reference_png_mask.py      # needs PIL and /root/reference (build container only)
"""
import os

import numpy as np
from PIL import Image

SRC = "/root/reference/teapot_4k_tris.png"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_png_mask.npz")

im = np.asarray(Image.open(SRC).convert("RGB")).astype(np.int32)
assert im.shape == (2160, 3840, 3)
sky = np.array([128, 178, 255])          # the sky constant of the revision that wrote the PNG (today: 128, 180, 255)
sub = im[4::8, 4::8]                     # 270 x 480 samples at pixel (8i+4, 8j+4)
is_hit = (np.abs(sub - sky).sum(-1) > 6)
np.savez_compressed(OUT, hit=np.packbits(is_hit), shape=np.array(is_hit.shape), step=np.array(8), offset=np.array(4),
                    source=np.array("teapot_4k_tris.png of gerikkub/rust_raytrace (3840x2160)"))
print("hit fraction", is_hit.mean(), "->", OUT, os.path.getsize(OUT), "bytes")
