"""Copy the reference files the diffusion path imports into the git-ignored baseline/_ref/ (run in the build
container, where /root/reference exists; `__graft_entry__.build()` calls this).  The copy travels to the GPU box
with the `gpurun` snapshot so that `bench.py --impl reference` and the `cpu_baseline` / `gpu_eager_baseline` legs
time the UNMODIFIED reference module there; it is never committed (.gitignore)."""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
FILES = [
    "models/__init__.py",
    "models/generative/__init__.py",
    "models/generative/diffusion/__init__.py",
    "models/generative/diffusion/ddpm.py",
    "models/modules/__init__.py",
    "models/modules/attend.py",
    "utils/__init__.py",
    "utils/lightning_utils.py",
]


def make_ref(src="/root/reference") -> bool:
    if not os.path.isfile(os.path.join(src, "models/generative/diffusion/ddpm.py")):
        return os.path.isfile(os.path.join(DST, "models/generative/diffusion/ddpm.py"))
    for rel in FILES:
        s, d = os.path.join(src, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        if os.path.isfile(s):
            shutil.copyfile(s, d)
        elif rel.endswith("__init__.py"):
            open(d, "a").close()           # namespace marker the reference leaves implicit
    return True


if __name__ == "__main__":
    ok = make_ref(*(sys.argv[1:2]))
    print("baseline/_ref", "ready" if ok else "unavailable")
