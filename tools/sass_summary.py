#!/usr/bin/env python
"""Counts of the SASS mnemonics that show which hardware paths the built library uses (cuobjdump -sass libsst_b200.so):
1-D TMA bulk copies (UBLKCP), 256-bit global loads (LDG.E...256), mbarrier waits (SYNCS), bulk prefetch, votes, async copies.
tools/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "suffix-array-searching_b200", "libsst_b200.so")
PATTERNS = {
    "UBLKCP (cp.async.bulk: 1-D TMA copies global<->shared)": r"\bUBLKCP",
    "LDG.E...256 (32-byte global loads, new on sm_100)": r"\bLDG\.E[.\w]*\.256",
    "LDG.E...128": r"\bLDG\.E[.\w]*\.128",
    "SYNCS (mbarrier arrive / try_wait)": r"\bSYNCS\.",
    "LDGSTS (cp.async 4/8/16-byte)": r"\bLDGSTS",
    "VOTE / VOTEU (ballots)": r"\bVOTEU?\.",
    "REDUX": r"\bREDUX",
    "MATCH": r"\bMATCH",
    "SHFL": r"\bSHFL\.",
    "ATOMS (shared-memory atomics)": r"\bATOMS",
    "UTMALDG / UTMASTG (tensor-map TMA: none expected, the data is 1-D)": r"\bUTMA(LDG|STG)",
    "UTCHMMA / UTCQMMA / tcgen05 MMA (none expected: no contraction on this path)": r"\bUTC\w*MMA",
    "HMMA / IMMA (none expected)": r"\b[HI]MMA",
}


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per_fn = collections.defaultdict(collections.Counter)
    fn = "?"
    total = collections.Counter()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            fn = re.sub(r"\(.*", "", name.replace("sst::(anonymous namespace)::", "").replace("void ", ""))
            continue
        for label, pat in PATTERNS.items():
            if re.search(pat, line):
                total[label] += 1
                per_fn[label][fn] += 1
    print(f"# SASS mnemonic counts of {os.path.relpath(LIB, ROOT)} (sm_100a), by tools/sass_summary.py")
    for label in PATTERNS:
        print(f"{total[label]:6d}  {label}")
        for f, c in per_fn[label].most_common(6):
            if label.startswith(("UBLKCP", "LDG.E...256", "SYNCS", "REDUX", "ATOMS")):
                print(f"          {c:5d}  {f[:110]}")


if __name__ == "__main__":
    main()
