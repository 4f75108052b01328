"""Count the Blackwell-specific SASS mnemonics per kernel of the built library (no GPU needed):

    python tools/sass_summary.py [mellon_b200/libmellon_b200.so] > profiles/sass_rNN_final.txt

UTCIMMA = tcgen05.mma kind::i8, UTCCP = tcgen05.cp, LDTM = tcgen05.ld, UBLKCP = cp.async.bulk (TMA bulk copy),
SYNCS = mbarrier operations, DMMA / DFMA = the FP64 tensor / FMA pipes (see /opt/skills/guides/B200_PROFILING.md)."""
import os
import re
import subprocess
import sys
from collections import Counter, defaultdict

MNEMONICS = ("UTCIMMA", "UTCCP", "LDTM", "UBLKCP", "SYNCS", "DMMA", "DFMA", "F2I", "I2F", "PRMT")


def summarise(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], check=True, capture_output=True, text=True).stdout
    per = defaultdict(Counter)
    total = Counter()
    name = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(anonymous namespace\)::", "", name)
            name = re.sub(r"\(.*", "", name).replace("void ", "")
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and name:
            op = m.group(1)
            per[name]["instructions"] += 1
            total["instructions"] += 1
            for k in MNEMONICS:
                if op == k or op.startswith(k + "."):
                    per[name][k] += 1
                    total[k] += 1
    return per, total


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mellon_b200",
                                                             "libmellon_b200.so")
    per, total = summarise(lib)
    print(f"# SASS mnemonic counts per kernel of {os.path.basename(lib)} (cuobjdump -sass, sm_100a); kernels without any of the "
          "counted mnemonics are left out")
    print("kernel,instructions," + ",".join(MNEMONICS))
    for name in sorted(per, key=lambda n: -per[n]["UTCIMMA"] * 10**6 - per[n]["instructions"]):
        c = per[name]
        if not any(c[k] for k in MNEMONICS):
            continue
        print(f'"{name}",{c["instructions"]},' + ",".join(str(c[k]) for k in MNEMONICS))
    print(f'"TOTAL ({len(per)} kernels)",{total["instructions"]},' + ",".join(str(total[k]) for k in MNEMONICS))


if __name__ == "__main__":
    main()
