"""In-tree build of the engine: nvcc -> pepr_b200/libpeprml.so (sm_100a only) and the raxmlHPC-compatible CLI shim."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpeprml.so")
CLI = os.path.join(HERE, "bin", "peprml")
SOURCES = ["model.cpp", "host.cpp", "kernels.cu", "newview_mma.cu", "branch_mma.cu", "fused_mma.cu", "parsimony.cu", "engine.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-fvisibility=hidden,-Wall", "-I", os.path.join(HERE, "..", "include")]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "peprml.h"), __file__]
    if force or _stale(LIB, deps):
        objs, jobs = [], []
        os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
        for s in SOURCES:
            o = os.path.join(HERE, "build", s + ".o")
            if force or _stale(o, deps):
                cmd = [NVCC] + ARCH + COMMON + ["-x", "cu", "-c", os.path.join(CSRC, s), "-o", o]
                if verbose:
                    print(" ".join(cmd))
                jobs.append((cmd, subprocess.Popen(cmd)))   # the translation units compile side by side
            objs.append(o)
        for cmd, job in jobs:
            if job.wait() != 0:
                raise subprocess.CalledProcessError(job.returncode, cmd)
        cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-ldl", "-lpthread"]
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
    cli_src = os.path.join(HERE, "cli", "peprml_main.cpp")
    if os.path.exists(cli_src) and (force or _stale(CLI, [cli_src, LIB])):
        os.makedirs(os.path.dirname(CLI), exist_ok=True)
        cmd = [NVCC] + ARCH + ["-O2", "-std=c++17", "-I", os.path.join(HERE, "..", "include"), cli_src, "-o", CLI,
                               "-L", HERE, "-lpeprml", "-Xlinker", "-rpath,$ORIGIN/.."]
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
    return LIB


def build_variant(name, defines, verbose=False):
    """Experiment aid: the engine compiled with extra -D flags into pepr_b200/variants/libpeprml_<name>.so (git-ignored; loaded
    instead of the product library when PEPRML_LIB points at it, see engine.lib)."""
    out_dir = os.path.join(HERE, "variants")
    obj_dir = os.path.join(HERE, "build", "var_" + name)
    os.makedirs(out_dir, exist_ok=True)
    os.makedirs(obj_dir, exist_ok=True)
    lib = os.path.join(out_dir, "libpeprml_%s.so" % name)
    objs, jobs = [], []
    for s in SOURCES:
        o = os.path.join(obj_dir, s + ".o")
        cmd = [NVCC] + ARCH + COMMON + ["-D" + d for d in defines] + ["-x", "cu", "-c", os.path.join(CSRC, s), "-o", o]
        if verbose:
            print(" ".join(cmd))
        jobs.append((cmd, subprocess.Popen(cmd)))
        objs.append(o)
    for cmd, job in jobs:
        if job.wait() != 0:
            raise subprocess.CalledProcessError(job.returncode, cmd)
    subprocess.run([NVCC] + ARCH + ["-shared", "-o", lib] + objs + ["-ldl", "-lpthread"], check=True)
    return lib


if __name__ == "__main__":
    if "--variant" in sys.argv:  # python -m pepr_b200.build --variant er1 PML_EARLY_RELEASE=1 ...
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], sys.argv[i + 2:], verbose=True))
        sys.exit(0)

    build(force="--force" in sys.argv, verbose=True)
