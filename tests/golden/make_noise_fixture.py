"""Regenerates mirror_maze_b200/assets/noiseTexture-2.rgba8.gz from the reference's PNG (run in the authoring
container only; /root/reference does not exist on the GPU box).

The reference embeds textures/noiseTexture-2.png (src/main.rs:354) and uploads its bitmap as a 512x512 RGBA8Unorm
texture (src/main.rs:667-695).  The fixture is Pillow's raw RGBA decode (alpha is 255 everywhere, so AppKit's
possible premultiplication is a no-op; SURVEY Appendix E caveat)."""
import gzip
import hashlib
import os
import sys

from PIL import Image

SRC = "/root/reference/textures/noiseTexture-2.png"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "mirror_maze_b200", "assets", "noiseTexture-2.rgba8.gz")

if __name__ == "__main__":
    im = Image.open(SRC).convert("RGBA")
    assert im.size == (512, 512), im.size
    raw = im.tobytes()
    assert raw[:4] == bytes([128, 128, 128, 255])
    with open(DST, "wb") as f:
        with gzip.GzipFile(fileobj=f, mode="wb", mtime=0, compresslevel=9) as g:
            g.write(raw)
    print("png sha256", hashlib.sha256(open(SRC, "rb").read()).hexdigest())
    print("raw sha256", hashlib.sha256(raw).hexdigest(), "bytes", len(raw), "->", os.path.getsize(DST))
