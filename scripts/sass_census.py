"""Per-kernel census of the SASS mnemonics that prove a Blackwell-native kernel (B200_PROFILING.md,
"What proves a Blackwell-native kernel"): UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG/UBLKCP =
TMA, UTCBAR = tcgen05.commit, SYNCS = mbarrier, REDG = red.global, HMMA = legacy mma.sync.  CPU-only:
    python scripts/sass_census.py > profiles/r01_sass_census.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "certifiedgpt_b200", "lib", "libcgpt.so")
FAMILIES = [("UTCMMA", r"^UTC[A-Z]*MMA"), ("LDTM", r"^LDTM"), ("STTM", r"^STTM"), ("UTMALDG", r"^UTMALDG"),
            ("UTMASTG", r"^UTMASTG"), ("UBLKCP", r"^UBLKCP"), ("UTMAPF", r"^UTMAPF"), ("UTCBAR", r"^UTCBAR"),
            ("SYNCS", r"^SYNCS"), ("REDG", r"^REDG"), ("HMMA", r"^HMMA"), ("LDGSTS", r"^LDGSTS"),
            ("MUFU.EX2", r"^MUFU\.EX2")]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = {}
    per = collections.OrderedDict()
    variants = collections.defaultdict(set)
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur is not None:
            op = m.group(1)
            per[cur]["_total"] += 1
            for fam, pat in FAMILIES:
                if re.match(pat, op):
                    per[cur][fam] += 1
                    variants[fam].add(op)
    try:
        names = subprocess.run(["cu++filt"] + list(per), capture_output=True, text=True, check=True).stdout.splitlines()
        demangle = dict(zip(per, names))
    except Exception:
        pass
    arch = re.search(r"arch = (sm_\w+)", sass)
    print(f"# cuobjdump -sass certifiedgpt_b200/lib/libcgpt.so  ({arch.group(1) if arch else '?'}; {len(per)} kernels)")
    print("# families: " + "; ".join(f"{f}: {', '.join(sorted(v))}" for f, v in variants.items()))
    cols = [f for f, _ in FAMILIES]
    print(f"{'instr':>7} " + " ".join(f"{c:>8}" for c in cols) + "  kernel")
    for k, c in per.items():
        name = demangle.get(k, k)
        name = re.sub(r"\((int|bool)\)", "", name)
        name = re.sub(r"\(.*", "", name)[:90]
        print(f"{c['_total']:>7} " + " ".join(f"{c[f] or '.':>8}" for f in cols) + f"  {name}")
    tc = [k for k, c in per.items() if c["UTCMMA"]]
    legacy = [k for k, c in per.items() if c["HMMA"]]
    print(f"# kernels issuing tcgen05.mma: {len(tc)}; kernels on the legacy mma.sync path: {len(legacy)} "
          f"(flash_attn_kernel / attention backward variants: Q-Former, tiny shapes, the row-major ViT A/B path, fine-tune backward)")


if __name__ == "__main__":
    sys.exit(main())
