"""Per-kernel SASS evidence (B200_PROFILING.md mnemonics): counts of tcgen05 / TMEM / TMA instructions and of the legacy
tensor-core path in every kernel of leann_b200/libleann_b200.so.  Writes profiles/<round>_sass_summary.md.
    python scripts/sass_summary.py r02
"""
import re
import subprocess
import sys
from collections import Counter, OrderedDict
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
SO = ROOT / "leann_b200" / "libleann_b200.so"
PAT = OrderedDict([("UTC*MMA (tcgen05.mma)", r"\bUTC[A-Z]*MMA"), ("LDTM (tcgen05.ld)", r"\bLDTM"), ("STTM (tcgen05.st)", r"\bSTTM"),
                   ("UTMALDG (TMA load)", r"\bUTMALDG"), ("UTMASTG (TMA store)", r"\bUTMASTG"), ("SYNCS (mbarrier)", r"\bSYNCS"),
                   ("HMMA (mma.sync)", r"\bHMMA"), ("LDGSTS (cp.async)", r"\bLDGSTS"), ("MUFU.EX2", r"\bMUFU\.EX2"), ("ATOM/RED", r"\b(ATOMG|ATOMS|RED)\b")])


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    sass = subprocess.run(["cuobjdump", "-sass", str(SO)], capture_output=True, text=True, check=True).stdout
    kernels = OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = Counter()
            continue
        if cur and re.search(r"/\*[0-9a-f]{4}\*/", line):
            kernels[cur]["instructions"] += 1
            for name, pat in PAT.items():
                if re.search(pat, line):
                    kernels[cur][name] += 1
    names = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    out = [f"# SASS summary of leann_b200/libleann_b200.so ({tag})", "",
           "`cuobjdump -sass leann_b200/libleann_b200.so`, instruction counts per kernel (scripts/sass_summary.py).", "",
           "| kernel | instr | " + " | ".join(PAT) + " |", "|---|---:|" + "---:|" * len(PAT)]
    for (mangled, c), dem in zip(kernels.items(), names):
        short = re.sub(r"\(anonymous namespace\)::", "", dem)
        short = re.sub(r"\(.*$", "", short).replace("lb2::", "")
        out.append(f"| `{short}` | {c['instructions']} | " + " | ".join(str(c[n]) if c[n] else "" for n in PAT) + " |")
    p = ROOT / "profiles" / f"{tag}_sass_summary.md"
    p.write_text("\n".join(out) + "\n")
    print(p.read_text())


if __name__ == "__main__":
    main()
