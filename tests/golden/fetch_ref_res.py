"""Copies the reference's REAL material libraries — and, for Sponza, the textures beside them — from /root/reference/res
into tests/golden/ref_res/, byte for byte (run in the build container; /root/reference does not exist on the GPU box, so
the copies are committed).  The .obj files next to them are Git-LFS pointers and are not copied: geometry is
the procedural stand-in of wgpu-pathtracer_b200/standins.py.

    python tests/golden/fetch_ref_res.py          # idempotent; prints a sha256 manifest (ref_res/MANIFEST.txt)
"""
import hashlib
import os
import shutil

SRC = "/root/reference/res"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_res")
# whole directories (small) and single files
DIRS = ["sponza"]
FILES = ["classroom/classroom.mtl", "san_miguel/san-miguel.mtl", "bistro/bistro.mtl", "dragon.mtl", "cornell_box.mtl",
         "cornell_box_pbr.mtl", "glass_cube.mtl", "cube.mtl", "knob/mitsuba.mtl"]


def main():
    copied = []
    for d in DIRS:
        for name in sorted(os.listdir(os.path.join(SRC, d))):
            if name.endswith(".obj"):
                continue                      # LFS pointer
            copied.append(os.path.join(d, name))
    copied += FILES
    lines = []
    for rel in copied:
        src, dst = os.path.join(SRC, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        lines.append(f"{hashlib.sha256(open(dst, 'rb').read()).hexdigest()}  {os.path.getsize(dst):8d}  {rel}")
    with open(os.path.join(DST, "MANIFEST.txt"), "w") as f:
        f.write("# sha256, bytes, path relative to /root/reference/res (reference commit 92a9a01b)\n" + "\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
