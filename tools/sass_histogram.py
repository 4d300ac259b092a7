"""Opcode histogram of the built objects (cuobjdump -sass): evidence that the hot kernels use the intended hardware
paths - DMMA (FP64 tensor pipe), UBLKCP (TMA bulk copies), SYNCS (mbarrier), UTCHMMA / UTCQMMA (tcgen05.mma),
LDTM / STTM (tensor memory), UTCBAR (tcgen05.commit), LDGSTS (cp.async).  Writes profiles/r02_sass_opcodes.md."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "cggp_b200", "build")
KEYS = ["DMMA", "DFMA", "DMUL", "DADD", "UBLKCP", "SYNCS", "UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "LDGSTS",
        "MUFU", "HMMA", "FFMA", "BAR", "ATOM", "RED"]


def histogram(obj):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    per_fn, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per_fn[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur:
            per_fn[cur][m.group(1)] += 1
    return per_fn


def demangle(name):
    try:
        return subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name
    except Exception:
        return name


def main():
    lines = ["# SASS opcode histogram per object (round 2)", "",
             "`python tools/sass_histogram.py` over `cggp_b200/build/*.o` (nvcc 12.9, `-gencode arch=compute_100a,"
             "code=sm_100a`).  Counts are static instruction counts summed over all kernels of the object; the second "
             "table lists the headline kernels.", "",
             "| object | kernels | " + " | ".join(KEYS) + " |", "|---|---|" + "---|" * len(KEYS)]
    picks = []
    for obj in sorted(os.listdir(BUILD)):
        if not obj.endswith(".o"):
            continue
        per_fn = histogram(os.path.join(BUILD, obj))
        tot = collections.Counter()
        for c in per_fn.values():
            tot.update(c)
        lines.append(f"| {obj} | {len(per_fn)} | " + " | ".join(str(tot.get(k, 0)) for k in KEYS) + " |")
        for fn, c in per_fn.items():
            if any(s in fn for s in ("kfu_pipe_kernelILi3ELi3ELi16ELi6", "kfu_pipe8_kernelILi3ELi3", "cg_tail_kernelId",
                                     "gram_contract_kernel", "dmma_gemm_nt_kernelILi8", "nearest_center_dmma")):
                picks.append((obj, fn, c))
    lines += ["", "## Headline kernels", "", "| object | kernel | " + " | ".join(KEYS) + " |",
              "|---|---|" + "---|" * len(KEYS)]
    seen = set()
    for obj, fn, c in picks:
        d = demangle(fn)
        short = re.sub(r"\(.*", "", d)[:110]
        if short in seen and "gram_contract" in short:
            continue
        seen.add(short)
        lines.append(f"| {obj} | `{short}` | " + " | ".join(str(c.get(k, 0)) for k in KEYS) + " |")
    path = os.path.join(ROOT, "profiles", "r02_sass_opcodes.md")
    open(path, "w").write("\n".join(lines) + "\n")
    print(path)


if __name__ == "__main__":
    sys.exit(main())
