"""SASS evidence: count the Blackwell-specific mnemonics per kernel family in the shipped library.

    python tools/sass_table.py > profiles/r02_sass.md

`cuobjdump -sass dlpm_b200/libdlpm_b200.so` (sm_100a).  UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), LDTM = tcgen05.ld,
UTMALDG / UTMASTG = cp.async.bulk.tensor load / store (TMA), UTMAPF = TMA prefetch (tensormap / L2), UTCBAR = tcgen05.commit,
SYNCS = mbarrier operations, HMMA = mma.sync (the attention kernel), MUFU.* = special-function unit, IMAD.WIDE = the 32x32->64
products of Philox.  The mnemonic names are those of B200_PROFILING.md."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "dlpm_b200", "libdlpm_b200.so")
MN = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTCBAR", "SYNCS", "HMMA", "MUFU", "IMAD.WIDE"]


def family(name):
    name = re.sub(r"^void ", "", name).replace("dlpm::", "")
    m = re.match(r"(k_conv_tc)<(\d+), (\d+), (\d+), (\d+), (true|false), (true|false)>", name)
    if m:
        kind = "XF (normalise on load)" if m.group(6) == "true" else ("POST (producer-side GroupNorm)" if m.group(7) == "true" else "plain")
        return "k_conv_tc, " + kind
    return re.sub(r"<.*", "", re.sub(r"\(.*", "", name))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
    chunks = re.split(r"Function : \S+", sass)[1:]
    agg = collections.OrderedDict()
    for name, body in zip(names, chunks):
        fam = family(name)
        a = agg.setdefault(fam, collections.Counter())
        a["kernels"] += 1
        a["instructions"] += len(re.findall(r"^\s+/\*[0-9a-f]{4}\*/", body, flags=re.M))
        for mn in MN:
            if mn == "UTCHMMA":
                a[mn] += len(re.findall(r"\bUTCHMMA\b(?!\.2CTA)", body))
            else:
                a[mn] += len(re.findall(r"\b" + re.escape(mn) + r"\b", body))
    size = os.path.getsize(LIB)
    print("# SASS evidence (round 2)\n")
    print("`cuobjdump -sass dlpm_b200/libdlpm_b200.so` (%d bytes, %d kernels, sm_100a), mnemonic counts per kernel family "
          "(`python tools/sass_table.py`).  Counts are static instructions over all template instantiations of the family.\n" % (size, len(chunks)))
    print("| kernel family | kernels | SASS instr | " + " | ".join(MN) + " |")
    print("|---|---:|---:|" + "---:|" * len(MN))
    tot = collections.Counter()
    for fam, a in agg.items():
        tot.update(a)
        print("| `%s` | %d | %d | " % (fam, a["kernels"], a["instructions"]) + " | ".join(str(a[m]) for m in MN) + " |")
    print("| **total** | %d | %d | " % (tot["kernels"], tot["instructions"]) + " | ".join(str(tot[m]) for m in MN) + " |")
    print("\nThe convolution path is tcgen05 / TMEM / TMA throughout (`UTCHMMA`, `LDTM`, `UTMALDG`, `UTMASTG`); no `HMMA` outside "
          "`k_attention_mma` (contraction dims of 16-64 are below a tcgen05 tile, DESIGN.md section 4); the streaming kernels carry the "
          "Philox `IMAD.WIDE` products and `MUFU` ops discussed in `r01_ncu_stream.md` / `r02_noise.md`.")


if __name__ == "__main__":
    main()
