"""Stages the UNMODIFIED reference package for use as the checker / CPU baseline (test infrastructure).

The reference (WagnerGroup/pyqmc) is pure Python: there is nothing to compile.  This recipe copies the
``*.py`` files of ``/root/reference/pyqmc`` -- byte for byte, from where they lie -- into the git-ignored
``oracle/_ref/pyqmc`` so that they travel to the GPU box with the snapshot (``/root/reference`` does not
exist there).  Nothing under ``oracle/_ref`` is product source and ``pyqmc_b200`` never imports it; it is
used by ``tests/`` (parity of the device objects under the reference's own drivers and test harness) and by
``bench.py``'s CPU arm (``--impl reference`` / ``cpu_baseline`` with ``kind: "reference"``).

Run by ``__graft_entry__.build()`` when ``/root/reference`` is present; idempotent."""
import filecmp
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCE = "/root/reference"
TARGET = os.path.join(HERE, "_ref")


def staged():
    return os.path.isfile(os.path.join(TARGET, "pyqmc", "api.py"))


def stage(source=SOURCE, target=TARGET):
    """Returns the number of files copied (0 when the staged copy is already identical)."""
    src_pkg = os.path.join(source, "pyqmc")
    if not os.path.isdir(src_pkg):
        return 0
    copied = 0
    for d, dirs, files in os.walk(src_pkg):
        dirs[:] = [x for x in dirs if x != "__pycache__"]
        rel = os.path.relpath(d, source)
        os.makedirs(os.path.join(target, rel), exist_ok=True)
        for f in files:
            if not f.endswith(".py"):
                continue
            s, t = os.path.join(d, f), os.path.join(target, rel, f)
            if not (os.path.exists(t) and filecmp.cmp(s, t, shallow=False)):
                shutil.copyfile(s, t)
                copied += 1
    lic = os.path.join(source, "LICENSE")
    if os.path.exists(lic):
        shutil.copyfile(lic, os.path.join(target, "LICENSE"))
    return copied


if __name__ == "__main__":
    print(f"staged {stage()} files into {TARGET}")
