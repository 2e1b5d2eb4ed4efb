#!/usr/bin/env python
"""Copy the triangle-soup fixtures the BASELINE.json configs name from the reference's testdata into
data/scenes/ (git-ignored; travels to the GPU box with the gpurun snapshot, where /root/reference does not
exist).  These are DATA files (uint32 count + float4 vertices), not reference source."""
import os
import shutil
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/testdata"
DST = os.path.join(REPO, "data", "scenes")
FILES = ["bunny.bin", "cryteksponza.bin", "bistro_ext_part1.bin", "bistro_ext_part2.bin", "lucy.bin", "xyzrgb_dragon.bin",
         "legocar.bin", "head.bin"]


def main(which=None):
    if not os.path.isdir(SRC):
        print(f"fetch_scenes: {SRC} absent; keeping whatever is in {DST}")
        return 0
    os.makedirs(DST, exist_ok=True)
    for f in which or FILES:
        s, d = os.path.join(SRC, f), os.path.join(DST, f)
        if os.path.isfile(s) and not (os.path.isfile(d) and os.path.getsize(d) == os.path.getsize(s)):
            shutil.copyfile(s, d)
            print("copied", f)
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:] or None))
