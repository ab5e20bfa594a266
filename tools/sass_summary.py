"""cuobjdump -sass of libb200ltx.so -> profiles/r2_sass_mnemonics.md: per kernel, how many tcgen05 / TMEM / TMA
instructions it contains (the mnemonics of /opt/skills/guides/B200_PROFILING.md).  No GPU needed.

  python tools/sass_summary.py            # rewrites profiles/r2_sass_mnemonics.md"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COLS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTCBAR", "MUFU.EX2", "FFMA2", "HMMA"]


def short(name):
    import subprocess
    d = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
    d = d.split("(")[0].replace("void ", "").replace("b200::", "")
    return d


def table(per):
    rows = ["| kernel | " + " | ".join(COLS) + " |", "|---|" + "---|" * len(COLS)]
    for name, ops in per.items():
        if not any(ops[c] for c in COLS):
            continue
        rows.append("| `" + short(name) + "` | " + " | ".join(str(ops[c]) if ops[c] else "" for c in COLS) + " |")
    total = {c: sum(ops[c] for ops in per.values()) for c in COLS}
    rows.append("| **all " + str(len(per)) + " kernels** | " + " | ".join(str(total[c]) for c in COLS) + " |")
    return "\n".join(rows)


def main():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, ROOT)
    import test_sass_evidence as t
    per = t.parse_sass()
    out = os.path.join(ROOT, "profiles", "r2_sass_mnemonics.md")
    with open(out, "w") as f:
        f.write("# SASS of the shipped library: tcgen05 / TMEM / TMA instruction counts per kernel\n\n"
                "`cuobjdump -sass video-generation-for-human-avatars_b200/libb200ltx.so`, parsed by `tools/sass_summary.py` "
                "(kept current by `tests/test_sass_evidence.py`).  `UTCHMMA` = tcgen05.mma (`.2CTA` = cta_group::2), `LDTM` / "
                "`STTM` = tcgen05.ld / st (TMEM), `UTMALDG` / `UTMASTG` / `UTMAREDG` = TMA load / store / reduce-add, `UTCBAR` = "
                "tcgen05.commit, `MUFU.EX2` = ex2.approx, `FFMA2` = packed fp32x2 FMA.  Kernels without any of them (the "
                "element-wise family) are left out of the rows; `HMMA` (the warp-level mma.sync path) is absent everywhere.\n\n")
        f.write(table(per) + "\n")
    print(open(out).read())


if __name__ == "__main__":
    main()
