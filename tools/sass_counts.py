"""Static SASS instruction counts of hippyflow_b200/libhfb200.so per kernel (cuobjdump -sass), as the markdown table of
profiles/r02_sass_counts.md.

    python tools/sass_counts.py > profiles/r02_sass_counts.md
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "hippyflow_b200", "libhfb200.so")
COLS = [("DMMA.8x8x4", r"\bDMMA\.8x8x4"), ("UTMALDG (TMA tile)", r"\bUTMALDG"), ("UBLKCP (TMA linear)", r"\bUBLKCP"),
        ("LDGSTS (cp.async)", r"\bLDGSTS"), ("SYNCS (mbarrier)", r"\bSYNCS"), ("USETMAXREG", r"\bUSETMAXREG"), ("DFMA", r"\bDFMA"),
        ("STG.256", r"\bSTG\.E(\.\w+)*\.256"), ("ld/st .STRONG.SYS", r"\b(LDG|STG|LD|ST)\.E(\.\w+)*\.STRONG\.SYS")]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True, check=True).stdout
    counts = collections.OrderedDict()
    inst = collections.Counter()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            dem = subprocess.run(["c++filt", m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip()
            name = re.sub(r"^void ", "", dem)
            name = re.sub(r"^hfb::", "", name)
            name = re.split(r"[<(]", name)[0]
            peer = ", true>" in dem and "dgemm_dmma_kernel" in dem
            cur = name + (" <PEER>" if peer else "")
            counts.setdefault(cur, collections.Counter())
            inst[cur] += 1
            continue
        if cur is None:
            continue
        for label, pat in COLS:
            if re.search(pat, line):
                counts[cur][label] += 1
    print("# Round 2 — SASS instruction counts of `hippyflow_b200/libhfb200.so` (sm_100a)\n")
    print("`cuobjdump -sass hippyflow_b200/libhfb200.so`, occurrences per kernel (static counts, `tools/sass_counts.py`). tcgen05 "
          "(`UTC*MMA`, `LDTM`) is absent by design: it has no f64 kind, so the FP64 tensor path of Blackwell is `DMMA.8x8x4` fed by "
          "TMA (`UTMALDG` tiled, `UBLKCP` linear) and `mbarrier` (`SYNCS`). `dgemm_dmma_kernel <PEER>` are the TN instantiations "
          "whose epilogue stores the tile with 256-bit `STG` into the owning rank's exchange buffer (a CUDA-IPC peer address: the "
          "fused lift + reduce-scatter); `peer_barrier_kernel` holds the system-scope release / acquire accesses of the exchange.\n")
    print("| kernel | " + " | ".join(c for c, _ in COLS) + " |")
    print("|---|" + "---|" * len(COLS))
    tot = collections.Counter()
    for k, c in counts.items():
        print("| `%s` (%d instantiation%s) | " % (k, inst[k], "" if inst[k] == 1 else "s") + " | ".join(str(c[l]) for l, _ in COLS) + " |")
        tot.update(c)
    print("| **total** | " + " | ".join(str(tot[l]) for l, _ in COLS) + " |")


if __name__ == "__main__":
    main()
