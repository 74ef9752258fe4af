"""SASS evidence per kernel family of eigencuda_b200/lib/libeigencuda_b200.so (cuobjdump -sass):
a mnemonic histogram per kernel and a short excerpt of the hot loop of one representative
instantiation per family.  Runs without a GPU.
    python tools/sass_excerpt.py > profiles/r02_sass_excerpts.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "eigencuda_b200", "lib", "libeigencuda_b200.so")
KEYS = ["UTMALDG", "UTMASTG", "UTMAPF", "DMMA", "DFMA", "UTCIMMA", "UTCBAR", "LDTM", "SYNCS", "LDS.128", "LDS",
        "STS", "STG.E.128", "STG", "LDG", "USETMAXREG", "PREEXIT", "ACQBULK", "BAR.SYNC", "ATOM", "RED", "MEMBAR", "CCTL"]

sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
funcs, name = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = m.group(1)
        funcs[name] = []
    elif name and re.match(r"\s*/\*[0-9a-f]{4,}\*/", line):
        funcs[name].append(line.rstrip())


def demangle(n):
    return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()


def mnemonic(line):
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    return m.group(1) if m else ""


print(f"# {os.path.relpath(LIB, ROOT)}: {len(funcs)} kernels\n")
print("## mnemonic counts per kernel (columns: " + " ".join(KEYS) + ")\n")
for n, body in funcs.items():
    ops = [mnemonic(l) for l in body]
    counts = []
    for k in KEYS:
        if k in ("LDS", "STG"):   # the plain forms: exclude the .128 variants counted separately
            counts.append(sum(1 for o in ops if o.startswith(k) and ".128" not in o))
        else:
            counts.append(sum(1 for o in ops if o.startswith(k)))
    print(f"{demangle(n)[:150]}\n    instructions {len(body):6d} | " + " ".join(f"{k}={c}" for k, c in zip(KEYS, counts) if c))
print()


def excerpt(pattern, key, before, after, title, last_cluster=False):
    for n, body in funcs.items():
        d = demangle(n)
        if re.search(pattern, d):
            idx = [i for i, l in enumerate(body) if mnemonic(l).startswith(key)]
            if not idx:
                continue
            if last_cluster:   # start of the last run of `key` instructions (gaps of at most 40 instructions)
                start = idx[-1]
                for a, b in zip(reversed(idx[:-1]), reversed(idx[1:])):
                    if b - a > 40:
                        break
                    start = a
                idx = [start]
            lo, hi = max(0, idx[0] - before), min(len(body), idx[0] + after)
            print(f"## {title}\n# {d[:200]}\n# first {key} at instruction {idx[0]} of {len(body)}")
            for l in body[lo:hi]:
                print(re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", l))
            print()
            return


excerpt(r"dgemm_tma_dmma_kernel<ec::TileCfg<128, 128.*false, false>", "UTMALDG", 12, 14, "FP64 hot kernel, producer: TMA loads signalled on mbarriers")
excerpt(r"dgemm_tma_dmma_kernel<ec::TileCfg<128, 128.*false, true>", "DMMA", 30, 24, "FP64 hot kernel (NN), fused tile turn-around: DMUL + STG.E.128 of the finished tile, then the new tile's DMMAs with RZ as C")
excerpt(r"dgemm_tma_dmma_kernel<ec::TileCfg<128, 128.*false, true>", "DMMA", 40, 30, "FP64 hot kernel (NN), consumer main loop: mbarrier wait, deferred release, LDS.128 fragments feeding DMMA.8x8x4", last_cluster=True)
excerpt(r"dgemm_tma_dmma_kernel<ec::TileCfg<128, 128.*false, false>", "PREEXIT", 2, 4, "FP64 hot kernel, programmatic dependent launch (griddepcontrol)")
excerpt(r"dgemm_tma_dmma_kernel<ec::TileCfg<128, 16", "DMMA", 10, 20, "128x16 tile (skinny products)")
excerpt(r"ozaki_gemm_kernel", "UTCIMMA", 10, 16, "Ozaki int8 backend: tcgen05.mma.kind::i8 (UTCIMMA)")
excerpt(r"ozaki_gemm_kernel", "LDTM", 6, 10, "Ozaki int8 backend: epilogue reads accumulators from TMEM (LDTM)")
excerpt(r"dgemm_simt_kernel", "DFMA", 6, 12, "generic fallback: FP64 FMA")
