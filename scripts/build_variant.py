"""Build a VARIANT of the library for an A/B run on the GPU box: the named translation units are recompiled with extra
nvcc flags (typically -D switches) and linked with the regular objects of everything else into
probdiffeq_b200/lib/variants/libprobdiffeq_b200_<name>.so (git-ignored, travels with the snapshot). Select it with
PDEQ_B200_LIB=<path>. usage: python scripts/build_variant.py <name> <unit.cu>[,<unit.cu>...] [nvcc flags...]"""
import pathlib, subprocess, sys

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
from probdiffeq_b200 import build as b  # noqa: E402

name, units, flags = sys.argv[1], sys.argv[2].split(","), sys.argv[3:]
b.build()
out_dir = b.ROOT / "lib" / "variants"
obj_dir = b.ROOT / "build" / ("variant_" + name)
out_dir.mkdir(parents=True, exist_ok=True)
obj_dir.mkdir(parents=True, exist_ok=True)
objs = []
for src in b.sources():
    if src.name in units:
        obj = obj_dir / (src.stem + ".o")
        cmd = [b._nvcc(), *b.NVCC_FLAGS, *flags, "-Xptxas=-v", "-c", str(src), "-o", str(obj)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise SystemExit(res.stderr)
        sys.stderr.write("\n".join(l for l in res.stderr.splitlines() if "registers" in l or "spill" in l) + "\n")
        objs.append(obj)
    else:
        objs.append(b.OBJ / (src.stem + ".o"))
lib = out_dir / f"libprobdiffeq_b200_{name}.so"
res = subprocess.run([b._nvcc(), "-shared", "-o", str(lib), *map(str, objs), "-ldl"], capture_output=True, text=True)
if res.returncode != 0:
    raise SystemExit(res.stderr)
print(lib)
