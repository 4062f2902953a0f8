"""Tuning aid for the forward attention kernel's FMA-pipe exponentials: builds one copy of the library per FDB_POLY_MASK
(which of the 16 pairs of every 32-column half use exp2_poly_x2 instead of MUFU.EX2) and times each at the BASELINE shape.
   python tools/micro/poly_exp_variants.py build            # here (nvcc cross-compiles)
   python tools/micro/poly_exp_variants.py run              # on the GPU box
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SRC = os.path.join(ROOT, "finetrainers_b200", "csrc")
MASKS = [0x0000, 0x8080, 0x8888, 0xA8A8, 0xAAAA]
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
         "-diag-suppress", "177", "-shared", "-cudart", "static"]


def lib_of(mask):
    return os.path.join(HERE, f"libb2d_poly_{mask:04x}.so")


if sys.argv[1] == "build":
    srcs = [os.path.join(SRC, f) for f in sorted(os.listdir(SRC)) if f.endswith(".cu")]
    procs = [subprocess.Popen(["nvcc", *FLAGS, f"-DFDB_POLY_MASK=0x{m:04x}u", "-o", lib_of(m), *srcs]) for m in MASKS]
    assert all(p.wait() == 0 for p in procs)
    print("built", [os.path.basename(lib_of(m)) for m in MASKS])
elif sys.argv[1] == "run":
    for m in MASKS:
        try:
            r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "attn_bench.py"), "--no-sdpa", "--lib", lib_of(m)],
                               capture_output=True, text=True, timeout=150)
        except subprocess.TimeoutExpired:
            print(f"mask {m:04x}: TIMEOUT (hang)", flush=True)
            continue
        print(f"mask {m:04x} ({bin(m).count('1')}/16 pairs on the FMA pipe):", " | ".join(line for line in r.stdout.strip().splitlines()[-2:]), r.stderr[-300:])
