#!/usr/bin/env python
"""Static SASS summary of libbetazero_b200.so (cuobjdump -sass; no GPU needed).

    python profiles/sass_summary.py            # writes profiles/sass_summary.txt and profiles/env_inst.json

Per kernel: instruction count and the opcodes that prove the Blackwell path (UTCHMMA = tcgen05.mma, LDTM =
tcgen05.ld, UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit, UTMALDG = TMA tensor load, SYNCS = mbarrier), and the
pipe mix (ALU: LOP3 / SHF / IADD3 / ISETP / SEL / LEA / PRMT / VIADD / VIMNMX ..., FMA: IMAD / FFMA / FMUL / FADD).
For the env kernels (straight-line bitboard code, 2 boards per thread and loop trip) it also derives the ALU-pipe
instructions per board that bench.py multiplies with the measured boards/s for the INT32-roofline fraction."""
from __future__ import annotations

import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "betazero_b200", "libbetazero_b200.so")
KEY = ("UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTCBAR", "SYNCS", "REDUX", "CREDUX", "SHFL", "VOTE",
       "MUFU", "LDG", "STG", "LDS", "STS", "BAR")
ALU = {"LOP3", "SHF", "IADD3", "ISETP", "SEL", "LEA", "PRMT", "VIADD", "VIMNMX", "IABS", "FSEL", "FSETP", "FMNMX", "PLOP3",
       "SGXT", "BMSK", "IMNMX", "MOV", "LOP", "IADD"}
FMA = {"IMAD", "FFMA", "FMUL", "FADD", "HFMA2", "HADD2", "HMUL2", "FADD2", "FMUL2", "FFMA2"}
XU = {"POPC", "FLO", "BREV", "MUFU", "I2F", "F2I", "I2FP", "F2FP"}


# EXECUTED instructions per board of the round-1 ncu capture (profiles/r1_env_kernels_raw.csv, 2^26 boards):
# smsp__inst_executed.sum * 32 / boards, and its ALU-pipe part = pipe_alu utilisation * 0.5 warp-inst/clk/SMSP *
# active cycles (= inst / issue_active).  The static count above also holds the ragged-tail path, so it is ~2 % higher.
NCU_R1 = {
    "legal_mask_kernel": {"ncu_inst_per_board": 216.6, "ncu_alu_pipe_inst_per_board": 216.6 * 0.8917 * 0.5 / 0.523},
    "step_first_legal_kernel": {"ncu_inst_per_board": 432.1, "ncu_alu_pipe_inst_per_board": 432.1 * 0.9354 * 0.5 / 0.556},
}


def demangle(names):
    try:
        out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True, check=True).stdout
        return out.strip().split("\n")
    except Exception:
        return names


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    funcs, cur = collections.OrderedDict(), None
    for ln in sass.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = funcs.setdefault(m.group(1), [])
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", ln)
        if m and cur is not None:
            ins = m.group(1).strip()
            ins = re.sub(r"^@!?U?P\d+\s+", "", ins)  # drop the predicate
            cur.append(ins.split()[0])
    names = demangle(list(funcs))
    lines, env = [], {}
    lines.append(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}: {len(funcs)} kernels")
    lines.append("# kernel | instructions | key opcodes | pipe mix (ALU / FMA / XU / other)")
    for (mangled, ops), name in zip(funcs.items(), names):
        short = re.sub(r"\(anonymous namespace\)::|bz::|\(int\)|\(bool\)|void ", "", name)
        short = re.sub(r"\(.*$", "", short)
        full = collections.Counter(ops)
        base = collections.Counter(o.split(".")[0] for o in ops)
        key = {}
        for k in KEY:
            n = full_count(full, k)
            if n:
                key[k] = n
        alu = sum(v for k, v in base.items() if k in ALU)
        fma = sum(v for k, v in base.items() if k in FMA)
        xu = sum(v for k, v in base.items() if k in XU)
        lines.append(f"{short} | {len(ops)} | " + ", ".join(f"{k} x{v}" for k, v in key.items()) +
                     f" | {alu} / {fma} / {xu} / {len(ops) - alu - fma - xu}")
        if re.match(r"(legal_mask_kernel|step_first_legal_kernel|apply_kernel|terminal_kernel)<true>", short):
            # the <true> instantiation is the 8x8 fast path: one loop trip handles 2 boards (one 128-bit load per array)
            k = short.split("<")[0]
            env[k] = {"alu_pipe_inst_per_board": alu / 2, "fma_pipe_inst_per_board": fma / 2,
                      "sass_inst_per_board": len(ops) / 2}
            if k in NCU_R1:
                env[k].update(NCU_R1[k])
    txt = "\n".join(lines) + "\n"
    open(os.path.join(ROOT, "profiles", "sass_summary.txt"), "w").write(txt)
    json.dump({"source": "static count: python profiles/sass_summary.py (cuobjdump -sass of the shipped library; whole kernel "
                         "body / 2 boards per loop trip, so prologue instructions are included: an upper bound within ~3 %)",
               "kernels": env}, open(os.path.join(ROOT, "profiles", "env_inst.json"), "w"), indent=1)
    sys.stdout.write(txt)


def full_count(counter, key):
    """instructions whose opcode is `key` or starts with `key.` (a key with a dot matches that suffix exactly)"""
    if "." in key:
        return sum(v for k, v in counter.items() if k == key or k.startswith(key + "."))
    return sum(v for k, v in counter.items() if k.split(".")[0] == key)


if __name__ == "__main__":
    main()
