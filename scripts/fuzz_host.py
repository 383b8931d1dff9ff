"""AddressSanitizer / UBSan fuzzing of the host-side parsers (COLLADA importer, TGA / BMP readers).

    python scripts/fuzz_host.py [seed] [iterations]

Builds tests/fuzz/{import,image}_driver.cpp with the host sources under -fsanitize=address,undefined into a scratch directory,
mutates the test fixture (truncation, byte flips, deleted / duplicated chunks, hostile numbers) and runs the drivers on every
mutation.  A driver touches every array the renderer would read.  Exit code 1 when a sanitizer reports anything.
Round 1: 1 900 COLLADA mutations found two out-of-bounds reads (negative <input offset>, negative <vcount>), fixed in
host_scene_import.cpp with a regression test in tests/test_scene_import.py; 900 image mutations found nothing.
"""
import os
import random
import re
import struct
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import collada_fixture as cf  # noqa: E402

CSRC = os.path.join(ROOT, "oppositerenderer_b200", "csrc")
SAN = ["g++", "-g", "-O1", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-std=c++17", "-I" + os.path.join(ROOT, "include")]


def build(work):
    subprocess.check_call(SAN + [os.path.join(ROOT, "tests", "fuzz", "import_driver.cpp")] +
                          [os.path.join(CSRC, f) for f in ("host_scene_import.cpp", "host_io.cpp", "host_api.cpp")] + ["-o", os.path.join(work, "import_driver")])
    subprocess.check_call(SAN + [os.path.join(ROOT, "tests", "fuzz", "image_driver.cpp"), os.path.join(CSRC, "host_io.cpp"), "-o", os.path.join(work, "image_driver")])


def run(driver, path):
    r = subprocess.run([driver, path], capture_output=True, text=True, errors="replace", timeout=120)
    bad = r.returncode != 0 or "ERROR" in r.stderr or "runtime error" in r.stderr
    return bad, "\n".join(l for l in r.stderr.splitlines() if "ERROR" in l or "runtime error" in l or re.search(r" #[0-3] ", l))[:900]


def mutate_text(rng, src):
    b = bytearray(src)
    mode = rng.randrange(6)
    if mode == 0:
        b = b[:rng.randrange(len(b))]
    elif mode == 1:
        for _ in range(rng.randrange(1, 20)):
            b[rng.randrange(len(b))] = rng.randrange(256)
    elif mode == 2:
        i = rng.randrange(len(b)); del b[i:min(len(b), i + rng.randrange(1, 400))]
    elif mode == 3:
        i = rng.randrange(len(b)); j = min(len(b), i + rng.randrange(1, 400)); b[i:i] = b[i:j]
    else:
        s = b.decode("latin1")
        for _ in range(rng.randrange(1, 4)):
            m = rng.choice(list(re.finditer(r"-?\d+", s)))
            s = s[:m.start()] + str(rng.choice([-1, 99999999, 0, 2 ** 31, 7, -5, 1000])) + s[m.end():]
        b = bytearray(s.encode("latin1"))
    return bytes(b)


def bmp(rng, w, h, bpp):
    row = (w * bpp // 8 + 3) & ~3
    data = bytes(rng.randrange(256) for _ in range(row * h))
    return b"BM" + struct.pack("<IHHI", 54 + len(data), 0, 0, 54) + struct.pack("<IiiHHIIiiII", 40, w, h, 1, bpp, 0, len(data), 2835, 2835, 0, 0) + data


def mutate_image(rng, src):
    b = bytearray(src)
    mode = rng.randrange(4)
    if mode == 0:
        b = b[:rng.randrange(len(b))]
    elif mode == 1:
        for _ in range(rng.randrange(1, 6)):
            b[rng.randrange(min(len(b), 60))] = rng.randrange(256)
    elif mode == 2:
        i = rng.randrange(min(len(b), 54)); b[i:i + 4] = struct.pack("<i", rng.choice([-1, 0, 2 ** 31 - 1, 70000, -70000, 1 << 20]))
    else:
        for _ in range(rng.randrange(1, 20)):
            b[rng.randrange(len(b))] = rng.randrange(256)
    return bytes(b)


def main():
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    rng = random.Random(seed)
    work = tempfile.mkdtemp(prefix="opr_fuzz_")
    build(work)
    dae, tex = cf.write_room(os.path.join(work, "room"))
    src = open(dae, "rb").read()
    images = [open(os.path.join(os.path.dirname(dae), "floor.tga"), "rb").read(), bmp(rng, 5, 3, 24), bmp(rng, 4, 4, 32)]
    bad = 0
    for it in range(iters):
        f = os.path.join(os.path.dirname(dae), "fuzz.dae")
        open(f, "wb").write(mutate_text(rng, src))
        b, msg = run(os.path.join(work, "import_driver"), f)
        if b:
            bad += 1; keep = os.path.join(work, "bad_%d.dae" % it); os.replace(f, keep); print("importer:", keep); print(msg)
        img = mutate_image(rng, rng.choice(images))
        f = os.path.join(work, "fuzz" + (".bmp" if img[:2] == b"BM" else ".tga"))
        open(f, "wb").write(img)
        b, msg = run(os.path.join(work, "image_driver"), f)
        if b:
            bad += 1; keep = os.path.join(work, "badimg_%d%s" % (it, f[-4:])); os.replace(f, keep); print("image reader:", keep); print(msg)
    print("seed %d, %d mutations each: %d sanitizer reports (scratch: %s)" % (seed, iters, bad, work))
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
