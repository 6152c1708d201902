"""Evidence that libcldrd.so carries Blackwell tensor-core / TMA / TMEM code: per-kernel SASS opcode histogram of the
in-tree library plus the MMA issue loop of the default scan kernel.
    python tools/sass_extract.py > profiles/r02_sass_extract.txt
(mnemonics: profiling recipe — UTCHMMA = tcgen05.mma kind::f16 / tf32 (UTCHMMA.2CTA = cta_group::2), UTMALDG = TMA tensor
load, LDTM = tcgen05.ld from TMEM, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "cl-drd_b200", "cldrd", "libcldrd.so")
KEY = ("UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "UTCCP", "SYNCS", "HMMA", "FFMA", "ATOM", "RED", "LDG", "STG",
       "ELECT", "UCGABAR", "MEMBAR")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    fn, hist, body = None, collections.OrderedDict(), collections.defaultdict(list)
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            fn = re.sub(r"\(.*", "", fn)
            hist[fn] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and fn:
            op = m.group(1)
            hist[fn][op] += 1
            body[fn].append(line.rstrip())
    print(f"# {os.path.relpath(LIB, ROOT)}: SASS opcode counts per kernel (sm_100a), grouped by mnemonic stem")
    for fn, c in hist.items():
        total = sum(c.values())
        groups = collections.Counter()
        for op, n in c.items():
            for k in KEY:
                if op.startswith(k):
                    groups[op if k.startswith(("UTC", "UTM", "LDTM")) else k] += n
                    break
        shown = ", ".join(f"{k}={v}" for k, v in sorted(groups.items()))
        print(f"{fn}: {total} instructions; {shown}")
    # the MMA issue loop of the default (cta_group::2, f16, filter) scan kernel
    target = next((f for f in body if "scan_tc2_kernel<0, 2>" in f), None)
    if target:
        lines = body[target]
        idx = [i for i, ln in enumerate(lines) if "UTCHMMA" in ln]
        if idx:
            lo, hi = max(0, idx[0] - 25), min(len(lines), idx[-1] + 12)
            print(f"\n# {target}: instructions {lo}..{hi} (the single-thread tcgen05.mma issue loop: descriptor updates, "
                  f"UTCHMMA.2CTA, commit)")
            for ln in lines[lo:hi]:
                print(re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", ln))
    return 0


if __name__ == "__main__":
    sys.exit(main())
