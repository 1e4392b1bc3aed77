"""Build libopd_b200.so (every CUDA source under csrc/) for sm_100a, in-tree.

The library is plain CUDA + a C ABI (include/opd_b200.h); it does not link against torch.
`python -m office_person_detection_vit_b200.build` or `__graft_entry__.build()` runs this.
"""

from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
LIB = PKG / "libopd_b200.so"
PROBE_LIB = PKG / "libopd_probe.so"      # measurement probes (csrc/probe/): a separate library, never loaded by the product path
OBJ_DIR = PKG / "build"

# measurement builds: OPD_EXTRA_NVCC_FLAGS="-DOPD_GEMM_PROBE -DOPD_STEM_PROBE" python -m office_person_detection_vit_b200.build
NVCC_FLAGS = os.environ.get("OPD_EXTRA_NVCC_FLAGS", "").split() + [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    f"-I{ROOT / 'include'}", f"-I{CSRC}", f"-I{CSRC / 'probe'}",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (cand == "nvcc" or Path(cand).exists()):
            return cand
    raise RuntimeError("nvcc not found")


def _digest(src: Path) -> str:
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    h.update(src.read_bytes())
    for hdr in sorted(list(CSRC.glob("*.h")) + list(CSRC.glob("*.cuh")) + list((CSRC / "probe").glob("*.h")) +
                      list((ROOT / "include").glob("*.h"))):
        h.update(hdr.read_bytes())
    return h.hexdigest()


def build(verbose: bool = False, force: bool = False) -> Path:
    """Compile changed sources to objects (parallel) and link the shared library."""
    OBJ_DIR.mkdir(exist_ok=True)
    nvcc = _nvcc()
    sources = sorted(CSRC.glob("*.cu"))
    if not sources:
        raise RuntimeError(f"no CUDA sources under {CSRC}")

    def compile_one(src: Path) -> tuple[Path, bool]:
        obj = OBJ_DIR / (src.stem + ".o")
        stamp = OBJ_DIR / (src.stem + ".sha")
        dig = _digest(src)
        if not force and obj.exists() and stamp.exists() and stamp.read_text() == dig:
            return obj, False
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        res = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src.name}")
        stamp.write_text(dig)
        return obj, True

    probes = sorted((CSRC / "probe").glob("*.cu"))
    with ThreadPoolExecutor(max_workers=min(8, len(sources) + len(probes))) as pool:
        results = list(pool.map(compile_one, sources + probes))
    n = len(sources)

    def link(lib: Path, res: list, extra: list[str]) -> None:
        if force or any(changed for _, changed in res) or not lib.exists():
            cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", str(lib), *[str(o) for o, _ in res],
                   *extra, "-lcudart", "-ldl"]
            out = subprocess.run(cmd, capture_output=True, text=True)
            if out.returncode != 0:
                sys.stderr.write(out.stdout + out.stderr)
                raise RuntimeError(f"link of {lib.name} failed")

    link(LIB, results[:n], [])
    # the probes use the product library's helpers (error string, tensor maps): linked against it, found next to it at run time
    link(PROBE_LIB, results[n:],
         [f"-L{PKG}", "-lopd_b200", "-Xlinker", "-rpath=$ORIGIN"])
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
