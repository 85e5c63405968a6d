"""SASS listings of the hot kernels, from the in-tree library, for profiles/ (python tools/dump_sass.py r02).
Each file starts with an opcode histogram (the mnemonics that prove which pipe a kernel is on: DMMA, UTCIMMA /
UTCMMA, LDTM, UBLKCP, LDGSTS, SYNCS) followed by the listing without the encoding columns."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "interpolater_b200", "lib", "libinterpolater_b200.so")
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
WANT = {
    "contract_idw": "_ZN3ipr15contract_kernelILi8ELb0ELb0ELi4EEEvNS_12ContractArgsE",
    "contract_numer": "_ZN3ipr15contract_kernelILi8ELb0ELb1ELi4EEEvNS_12ContractArgsE",
    "cressman": "_ZN3ipr15cressman_kernelILi8ELi2EEEvNS_12CressmanArgsE",
    "denom_umma": "_ZN3ipr17denom_umma_kernelENS_9DenomArgsE",
}
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
blocks = re.split(r"\n\s*Function : ", sass)
for short, mangled in WANT.items():
    body = next((b for b in blocks if b.startswith(mangled)), None)
    if body is None:
        print("missing:", mangled)
        continue
    lines = []
    hist = collections.Counter()
    for l in body.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s*/\*", l)
        if not m:
            continue
        ins = m.group(2).strip()
        lines.append(f"/*{m.group(1)}*/  {ins}")
        op = re.sub(r"^@!?U?P\d+\s+", "", ins).split()[0].split(".")[0]
        hist[op] += 1
    path = os.path.join(ROOT, "profiles", f"{tag}_sass_{short}.txt")
    with open(path, "w") as f:
        f.write(f"# {mangled}\n# cuobjdump -sass of interpolater_b200/lib/libinterpolater_b200.so (sm_100a), {len(lines)} instructions\n")
        f.write("# opcode histogram: " + ", ".join(f"{k} {v}" for k, v in hist.most_common()) + "\n\n")
        f.write("\n".join(lines) + "\n")
    key = {k: hist[k] for k in ("DMMA", "UTCIMMA", "UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UBLKCP", "UTMALDG", "LDGSTS", "SYNCS", "IMMA") if hist[k]}
    print(short, len(lines), "instructions;", key)
