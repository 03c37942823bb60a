"""Two frames of the reference's own rendering, decoded from /root/reference/out_files/1.gif (recorded by the reference's pyglet
viewer, gym_fortattack/fortattack.py:368-596, 700 x 700): frame 0 (the initial state: headings known, no lasers, no halos) and
frame 8 (fort disc unobstructed).  Stored as palette indices + palette.  Run in the build container (needs PIL):
    python tests/golden/make_render_golden.py"""
import os

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("FA_REFERENCE_DIR", "/root/reference")


def main():
    im = Image.open(os.path.join(REF, "out_files", "1.gif"))
    out = {"n_frames": np.array(im.n_frames), "size": np.array(im.size)}
    for k in (0, 8):
        im.seek(k)
        rgb = np.array(im.convert("RGB"))
        pal, idx = np.unique(rgb.reshape(-1, 3), axis=0, return_inverse=True)
        assert len(pal) < 256
        out["frame%d/palette" % k] = pal.astype(np.uint8)
        out["frame%d/index" % k] = idx.reshape(rgb.shape[:2]).astype(np.uint8)
    np.savez_compressed(os.path.join(HERE, "render_ref_frames.npz"), **out)
    print("wrote render_ref_frames.npz", os.path.getsize(os.path.join(HERE, "render_ref_frames.npz")), "bytes")


if __name__ == "__main__":
    main()
