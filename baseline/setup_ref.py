#!/usr/bin/env python
"""Create the reference sandbox baseline/_ref/ (git-ignored; it travels to the GPU box with gpurun):
an unmodified copy of the mounted reference tree with the exec bit set on its bundled engine.
Reference sources are never committed to this repository.

    python baseline/setup_ref.py [/root/reference]
"""
import os
import shutil
import stat
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    dst = os.path.join(HERE, "_ref")
    if not os.path.isdir(src):
        print("reference not mounted at %s; keeping %s as is" % (src, dst))
        return 0 if os.path.isdir(dst) else 1
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    shutil.copytree(src, dst, symlinks=True,
                    ignore=shutil.ignore_patterns("__pycache__", "*.pyc", ".git", "build", "*.so"))
    for root, dirs, files in os.walk(dst):     # the mount is read-only; make the copy writable
        for d in dirs:
            os.chmod(os.path.join(root, d), 0o755)
        for f in files:
            os.chmod(os.path.join(root, f), 0o644)
    jf = os.path.join(dst, "library", "jellyfish-linux")
    os.chmod(jf, os.stat(jf).st_mode | stat.S_IXUSR | stat.S_IXGRP | stat.S_IXOTH)
    print("reference sandbox ready:", dst)
    return 0


if __name__ == "__main__":
    sys.exit(main())
