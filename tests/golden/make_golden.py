"""Regenerates the committed golden fixtures from the reference checkout (run in the build
container, where /root/reference exists; the GPU box only ever reads the outputs).

  reference_hashes.json  the blockhash256 lines of lib/zosimos/tests/reference/*.crc.png
                         (one list per golden; several lines = accepted device variants)
  fixtures.npz           the two input images of lib/zosimos/tests/input/ decoded to RGBA8
                         arrays (the pixels the reference's tests feed to Pool::insert_srgb)

Usage: python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np
from PIL import Image

REF = "/root/reference/lib/zosimos/tests"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    if not os.path.isdir(REF):
        sys.exit("reference checkout not present")
    hashes = {}
    for name in sorted(os.listdir(os.path.join(REF, "reference"))):
        lines = open(os.path.join(REF, "reference", name)).read().split()
        key = name[: -len(".crc.png")]
        if all(len(l) == 64 for l in lines):  # derived.crc.png is a stale decimal CRC, skip
            hashes[key] = lines
    with open(os.path.join(HERE, "reference_hashes.json"), "w") as f:
        json.dump(hashes, f, indent=1, sort_keys=True)
    arrays = {}
    for name in ("background", "foreground"):
        arrays[name] = np.asarray(Image.open(os.path.join(REF, "input", name + ".png")).convert("RGBA"))
    np.savez_compressed(os.path.join(HERE, "fixtures.npz"), **arrays)
    print({k: v.shape for k, v in arrays.items()}, len(hashes), "goldens")


if __name__ == "__main__":
    main()
