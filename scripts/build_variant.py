#!/usr/bin/env python
"""Development tool: build an A/B variant of the library with extra nvcc flags (usually -D tuning macros).

    python scripts/build_variant.py bwd16 -DEGS_BWD_MIN_CTAS=16
    EGS_RASTER_LIB=easy_gaussian_splatting_b200/_C/variants/bwd16.so python bench.py ...

The variant .so lands in easy_gaussian_splatting_b200/_C/variants/ (git-ignored, travels to the GPU box)."""
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from easy_gaussian_splatting_b200 import build as B  # noqa: E402


def main():
    name, extra = sys.argv[1], sys.argv[2:]
    # --replace blend.cu=/path/to/other.cu : compile another file in place of a translation unit (A/B of a rewrite)
    replace = {}
    for a in [a for a in extra if a.startswith("--replace=")]:
        k, v = a[len("--replace="):].split("=", 1)
        replace[k] = Path(v)
        extra.remove(a)
    out_dir = B.OUT_DIR / "variants"
    obj_dir = out_dir / f"obj_{name}"
    obj_dir.mkdir(parents=True, exist_ok=True)
    nvcc = B._nvcc()

    def one(src):
        obj = obj_dir / (src.stem + ".o")
        real = replace.get(src.name, src)
        cmd = [nvcc, *B.ARCH, *B.COMMON, f"-I{B.CSRC}", *B.FLAGS.get(src.name, []), *extra, "-c", str(real), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(r.stdout + r.stderr)
        return obj, r.stdout + r.stderr

    with ThreadPoolExecutor(8) as ex:
        res = list(ex.map(one, B._sources()))
    lib = out_dir / f"{name}.so"
    subprocess.run([nvcc, *B.ARCH, "-shared", "-o", str(lib), *[str(o) for o, _ in res], "-lcudart"], check=True)
    (out_dir / f"{name}.ptxas.log").write_text("\n".join(t for _, t in res))
    for _, t in res:
        lines = t.splitlines()
        for i, ln in enumerate(lines):
            if "rasterize_" in ln and "Function properties" in ln:
                print(ln.split("for ")[-1][:60], "|", lines[i + 1].strip(), "|", lines[i + 2].strip())
    print(lib)


if __name__ == "__main__":
    main()
