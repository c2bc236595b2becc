#!/usr/bin/env python3
"""Copy the scene DATA the reference ships (scenes/coffee: 19 OBJ files + coffee.scene, MIT) from
/root/reference into scenes/coffee/.  The directory is git-ignored (15.9 MB of third-party data
stays out of history) but travels to the GPU box with the snapshot.  No source code is copied."""
import os
import shutil
import sys

SRC = "/root/reference/MinimalOptiX/scenes/coffee"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "scenes", "coffee")


def main():
    if not os.path.isdir(SRC):
        print("reference scenes not present; nothing to fetch")
        return 0
    os.makedirs(DST, exist_ok=True)
    n = 0
    for name in sorted(os.listdir(SRC)):
        if name.endswith(".obj") or name.endswith(".scene"):
            dst = os.path.join(DST, name)
            if not os.path.exists(dst) or os.path.getsize(dst) != os.path.getsize(os.path.join(SRC, name)):
                shutil.copyfile(os.path.join(SRC, name), dst)
            n += 1
    print(f"scenes/coffee: {n} files")
    return 0


if __name__ == "__main__":
    sys.exit(main())
