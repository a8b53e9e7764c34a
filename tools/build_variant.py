"""Builds liblwfb90 with extra compiler flags into build_variants/<name>.so (select it with LWFB_LIB=<path>).
usage: build_variant.py name [-DFLAG ...]      e.g.  build_variant.py free -DLWFB_LANES_FREE_RUNNING"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import importlib
b = importlib.import_module("lwfbrook90.jl_b200.build")


def build(name, flags, verbose=False):
    out_dir = os.path.join(ROOT, "build_variants")
    obj_dir = os.path.join("/tmp", "lwfb_variant_obj_" + name)
    os.makedirs(obj_dir, exist_ok=True)
    nvcc = b.nvcc_path()
    procs, objs = [], []
    for src, uf in b.UNITS:
        obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [nvcc] + b.NVCC_FLAGS + list(uf) + list(flags) + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
        procs.append((cmd, subprocess.Popen(cmd, cwd=b._CSRC)))
    for cmd, p in procs:
        if p.wait() != 0:
            raise subprocess.CalledProcessError(p.returncode, cmd)
    out = os.path.join(out_dir, name + ".so")
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", out] + objs, cwd=b._CSRC)
    return out


if __name__ == "__main__":
    print(build(sys.argv[1], sys.argv[2:], verbose=os.environ.get("VERBOSE") == "1"))
