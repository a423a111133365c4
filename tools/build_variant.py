"""Compile one translation unit with extra defines and link it with the objects of the regular build into
pantea_b200/variants/lib_<name>.so (selected at run time with PANTEA_B200_LIB).
usage: python tools/build_variant.py <name> <source.cu> "<-Dflags>" """
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from pantea_b200.csrc import build as B  # noqa: E402

name, src, flags = sys.argv[1], sys.argv[2], sys.argv[3].split()
B.build()
out_dir = ROOT / "pantea_b200" / "variants"
out_dir.mkdir(exist_ok=True)
obj = out_dir / f"{name}_{src}.o"
common = [B._nvcc(), *B.ARCH, "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", f"-I{ROOT / 'include'}", f"-I{B.HERE}"]
subprocess.run(common + flags + ["-c", str(B.HERE / src), "-o", str(obj)], check=True)
objs = [str(obj) if s == src else str(B.HERE / "build" / (s + ".o")) for s in B.SOURCES]
lib = out_dir / f"lib_{name}.so"
subprocess.run([B._nvcc(), *B.ARCH, "-shared", "-o", str(lib), *objs], check=True)
print(lib)
