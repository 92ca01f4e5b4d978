#!/usr/bin/env python3
"""Opcode histogram of the SASS of libsvi_gpu.so (sm_100a), in total and per kernel for the mnemonics that prove the
design: UTMALDG (TMA tile loads), SYNCS (mbarrier), HSET2 (packed fp16 compares of the BRIEF tests), POPC, REDUX (warp
arg-min), DADD (fp64 box sums), LDS/STS, BAR, ATOM*.   usage: python tools/sass_opcodes.py [lib] > profiles/r2_sass_opcodes.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "svi_mapper_b200/libsvi_gpu.so"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
tot, per, fn = collections.Counter(), collections.defaultdict(collections.Counter), None
arch = set(re.findall(r"arch = (\S+)", txt))
for line in txt.splitlines():
    m = re.match(r"\s+Function : (\S+)", line)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0].replace("void ", "")
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and fn:
        tot[m.group(1)] += 1
        per[fn][m.group(1)] += 1
KEY = ["UTMALDG", "SYNCS", "HSET2", "POPC", "REDUX", "DADD", "LDS", "STS", "LDG", "STG", "BAR", "ATOMS", "ATOMG", "RED", "SHFL", "F2F", "I2FP", "FMNMX", "PRMT", "LOP3"]
print(f"# cuobjdump -sass {lib}: arch {sorted(arch)}; {sum(tot.values())} instructions in {len(per)} kernels")
print("## whole library, 40 most frequent opcodes")
for op, n in tot.most_common(40):
    print(f"{n:8d} {op}")
print("## tensor-core / TMEM opcodes (none expected: nothing on this path is a dense contraction): " +
      (", ".join(f"{k}={v}" for k, v in tot.items() if re.match(r"UTC.*MMA|LDTM|STTM|HMMA|HGMMA", k)) or "none"))
print("## per kernel: " + " ".join(KEY))
for f in sorted(per):
    print(f"{f[:70]:70s} " + " ".join(f"{k}={per[f][k]}" for k in KEY if per[f][k]))
