#!/usr/bin/env python
"""Writes tests/golden/jpeg/*.jpg and the pixels libjpeg(-turbo) decodes them to (*.rgb, raw 8-bit
RGB rows top-down), using Pillow in THIS container.  QImage, which the reference loads textures with
(MinimalOptiX.cpp:445-479), sits on the same libjpeg arithmetic, so these are the golden vectors of the
JPEG branch of the texture path.  Run once; the outputs are committed."""
import io, json, os
import numpy as np
from PIL import Image

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "jpeg")


def picture(w, h, seed):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w].astype(np.float64)
    img = np.stack([127 + 120 * np.sin(x / 5.0 + seed) * np.cos(y / 7.0),
                    255 * x / max(1, w - 1),
                    255 * ((x // 4 + y // 4) % 2)], axis=-1)
    img += rng.normal(0, 12, img.shape)
    return np.clip(img, 0, 255).astype(np.uint8)


CASES = [
    # name, w, h, mode, save kwargs
    ("ycc420_q85", 37, 29, "RGB", dict(quality=85, subsampling="4:2:0")),
    ("ycc444_q95", 16, 16, "RGB", dict(quality=95, subsampling="4:4:4")),
    ("ycc422_q75", 33, 17, "RGB", dict(quality=75, subsampling="4:2:2")),
    ("grey_q80", 23, 31, "L", dict(quality=80)),
    ("prog420_q70", 40, 40, "RGB", dict(quality=70, subsampling="4:2:0", progressive=True)),
    ("prog444_q90", 19, 21, "RGB", dict(quality=90, subsampling="4:4:4", progressive=True)),
    ("prog_grey", 30, 14, "L", dict(quality=60, progressive=True)),
    ("opt420_q50", 64, 48, "RGB", dict(quality=50, subsampling="4:2:0", optimize=True)),
    ("rst420", 50, 34, "RGB", dict(quality=85, subsampling="4:2:0", restart_marker_blocks=3)),
    ("rst_prog422", 41, 23, "RGB", dict(quality=80, subsampling="4:2:2", progressive=True, restart_marker_rows=1)),
    ("tiny_3x3", 3, 3, "RGB", dict(quality=90, subsampling="4:2:0")),
    ("tiny_1x1", 1, 1, "RGB", dict(quality=90, subsampling="4:2:0")),
    ("tiny_5x2_422", 5, 2, "RGB", dict(quality=90, subsampling="4:2:2")),
    ("q10_420", 48, 40, "RGB", dict(quality=10, subsampling="4:2:0")),
    ("q100_444", 24, 24, "RGB", dict(quality=100, subsampling="4:4:4")),
    ("big420", 160, 120, "RGB", dict(quality=88, subsampling="4:2:0")),
]


def main():
    os.makedirs(OUT, exist_ok=True)
    index = []
    for i, (name, w, h, mode, kw) in enumerate(CASES):
        src = picture(w, h, i + 1)
        im = Image.fromarray(src if mode == "RGB" else src[..., 0], mode)
        buf = io.BytesIO()
        im.save(buf, "JPEG", **kw)
        data = buf.getvalue()
        dec = np.asarray(Image.open(io.BytesIO(data)).convert("RGB"))
        open(os.path.join(OUT, name + ".jpg"), "wb").write(data)
        dec.tofile(os.path.join(OUT, name + ".rgb"))
        index.append(dict(name=name, width=w, height=h, bytes=len(data)))
    json.dump(dict(generator="Pillow %s (bundled libjpeg-turbo)" % Image.__version__, cases=index),
              open(os.path.join(OUT, "index.json"), "w"), indent=1)
    print("wrote", len(index), "cases to", OUT)


if __name__ == "__main__":
    import PIL
    Image.__version__ = PIL.__version__
    main()
