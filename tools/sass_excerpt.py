#!/usr/bin/env python
"""profiles/r02_sass_excerpt.md: mnemonic histograms and the key instruction sequences of the matching kernels, from
`cuobjdump -sass` of the built library.   python tools/sass_excerpt.py > profiles/r02_sass_excerpt.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "linemod_pose_estimation_b200", "liblinemod_b200.so")


def functions():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    cur, fns = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            fns[cur] = []
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and cur:
            fns[cur].append((m.group(1), m.group(2).strip()))
    return fns


def mnemonic(ins):
    parts = ins.split()
    op = parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]
    return op


def histogram(body, n=22):
    c = collections.Counter(mnemonic(i).split(".")[0] for _, i in body)
    return c.most_common(n)


def forms(body, pats):
    c = collections.Counter()
    for _, i in body:
        op = mnemonic(i)
        if any(op.startswith(p) for p in pats):
            c[op] += 1
    return ", ".join("`%s` x%d" % kv for kv in sorted(c.items()))


def excerpt(body, start_pat, before, after, nth=0):
    hits = [k for k, (_, i) in enumerate(body) if re.search(start_pat, i)]
    if len(hits) <= nth:
        return "(pattern %s not found)" % start_pat
    k = hits[nth]
    return "\n".join("/*%s*/  %s ;" % a for a in body[max(0, k - before):k + after])


def main():
    fns = functions()
    pick = lambda key: next((v for k, v in fns.items() if key in k), [])
    print("# SASS of the production matching kernels (sm_100a, `cuobjdump -sass liblinemod_b200.so`, round-2 final build)\n")
    print("Made by `tools/sass_excerpt.py`.  Tensor-core and TMA-tensor mnemonics (`UTC*MMA`, `LDTM`, `UTMALDG`) are absent by design (a gather-"
          "accumulate, not a contraction); `UBLKCP` + `SYNCS` are the bulk-asynchronous copy of the tile records and its mbarrier.\n")
    for name, key in (("k_similarity_coarse_rec63 (u8-only coarse kernel, 80 registers, 3 CTAs/SM)", "k_similarity_coarse_rec63E"),
                      ("k_similarity_coarse_rec (general coarse kernel, u16 totals, 2 CTAs/SM)", "k_similarity_coarse_recE"),
                      ("k_refine_nib", "k_refine_nib")):
        body = pick(key)
        print("\n## `%s` -- %d SASS instructions (%.1f KB)\n" % (name, len(body), len(body) * 16 / 1024.0))
        print("| mnemonic | count |\n|---|---|")
        for k, v in histogram(body):
            print("| %s | %d |" % (k, v))
        print("\nMemory / synchronisation forms: " + forms(body, ("LDG", "LDS", "STG", "STS", "ATOM", "RED", "SYNCS", "UBLKCP", "SHFL", "SHF.R.W", "VIMNMX", "CREDUX", "VOTE")) + "\n")
    c63 = pick("k_similarity_coarse_rec63E")
    print("\n## Tile-record staging in the coarse kernel (arm the barrier, issue the bulk copy, wait on the phase)\n\n```")
    print(excerpt(c63, r"UBLKCP", 6, 3))
    print("...")
    print(excerpt(c63, r"SYNCS.PHASECHK", 1, 3))
    print("```\n")
    print("## A batch of six features in the coarse kernel: one `LOP3` + `IADD3` + `IMAD.X` per window address (the lane's plane pointer is "
          "pinned in a register pair), twelve loads in flight (`LDG.E.128.CONSTANT` + the following words), then `SHF.R.W` realignment, "
          "nibble sums and the even / odd split into u8 sums\n\n```")
    print(excerpt(c63, r"LDG\.E\.128\.CONSTANT", 14, 60, nth=6))
    print("```\n")
    rf = pick("k_refine_nib")
    print("## A step of the refinement on column-blocked planes: `LDS.64` of a feature's (offset0 | shift, offset1), two `LDG.E.64.CONSTANT` "
          "(the 16 rows of a chunk are 128 contiguous bytes), selects by the shift's high bit, two `SHF.R.W`, nibble sums 3 + 3 + 2\n\n```")
    hits = [k for k, (_, i) in enumerate(rf) if re.match(r"LDS\.64", i) and not i.startswith("@")]
    if hits:
        k = hits[0]
        print("\n".join("/*%s*/  %s ;" % a for a in rf[max(0, k - 2):k + 70]))
    print("```")


if __name__ == "__main__":
    sys.exit(main())
