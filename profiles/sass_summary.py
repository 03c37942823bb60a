#!/usr/bin/env python
"""Instruction-count summary of the product library's SASS (cuobjdump -sass), per kernel: the mnemonics that prove the
tensor-core / tensor-memory / bulk-copy paths (UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit,
UBLKCP = cp.async.bulk, SYNCS = mbarrier ops) and the warp-level primitives of the step kernels (VOTE, SHFL, BAR).
    python profiles/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "emergent-multiagent-strategies_b200", "libfortattack_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
KEYS = ("UTCHMMA", "UTCQMMA", "UTCOMMA", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "HMMA", "VOTE", "SHFL",
        "BAR", "WARPSYNC", "MUFU", "LDS", "STS", "LDG", "STG", "LDGSTS", "DFMA", "DMUL", "DADD")
cur, counts, total = None, collections.OrderedDict(), collections.Counter()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur)
        counts[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        op = m.group(1)
        counts[cur]["_all"] += 1
        for k in KEYS:
            if op.startswith(k):
                counts[cur][k] += 1
                total[k] += 1
                break
print("# cuobjdump -sass %s (sm_100a), instruction counts per kernel" % os.path.relpath(so, ROOT))
print("# totals: " + ", ".join("%s %d" % (k, total[k]) for k in KEYS if total[k]))
pat = sys.argv[1] if len(sys.argv) > 1 else r"mp_policy|tg_|fa_step_group_kernel<3, 3|fa_step_wide_kernel<3, 3|fa_step_kernel<3, 3|fa_step_group_kernel<5, 5|rl::|attn"
for name, c in counts.items():
    if re.search(pat, name):
        print("%-90s %6d instr  %s" % (name[:90], c["_all"], " ".join("%s=%d" % (k, c[k]) for k in KEYS if c[k])))
