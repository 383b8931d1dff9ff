"""Tiny PNG writer (no PIL in the image) + display transform of the reference GUI
(Gui/gui/RenderWidget.cpp:117-121,197: divide by iterations, pow(1/2.2))."""
import struct
import zlib

import numpy as np


def write_png(path, rgb8):
    h, w, _ = rgb8.shape
    raw = b"".join(b"\x00" + rgb8[y].tobytes() for y in range(h))

    def chunk(t, d):
        c = struct.pack(">I", len(d)) + t + d
        return c + struct.pack(">I", zlib.crc32(t + d) & 0xFFFFFFFF)

    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)) +
                chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))


def display(sum_buffer, iterations, gamma=2.2):
    img = np.clip(sum_buffer / float(iterations), 0, None) ** (1.0 / gamma)
    img = np.clip(img, 0, 1)[::-1]  # launch index y grows upward on screen
    return (img * 255 + 0.5).astype(np.uint8)
