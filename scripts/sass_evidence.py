#!/usr/bin/env python
"""Which kernels of libcmh.so contain tcgen05 / TMA / TMEM / NVLS instructions (SASS mnemonics, B200_PROFILING.md):
UTCIMMA = tcgen05.mma.kind::i8, UTCHMMA = kind::f16, UTMALDG = TMA tensor load, UBLKCP = bulk copy (1-D TMA), LDTM = tcgen05.ld,
UTCBAR = tcgen05.commit, VIMNMX3.S16x2 = packed 16-bit max (collect pass), LDGMC = multimem.ld_reduce (NVSwitch multicast
load-reduce); multimem.st has no mnemonic of its own: it is an STG.E.*.STRONG.SYS whose address is the multicast mapping."""
import collections
import os
import re
import subprocess
import sys

so = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "clip_based_cross_modal_hash_b200", "libcmh.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
pat = re.compile(r"\b(UTCIMMA|UTCHMMA|UTMALDG|UBLKCP|LDTM|UTCBAR|VIMNMX3\.S16x2|LDGMC|STG\.E\.(?:64|128)\.STRONG\.SYS|HMMA|POPC)[\w.]*")
per = collections.OrderedDict()
fn = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        per[fn] = collections.Counter()
        continue
    if fn:
        m = pat.search(line)
        if m:
            per[fn][m.group(0)] += 1
names = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
for raw, name in zip(per, names):
    if not per[raw]:
        continue
    short = re.sub(r"\(anonymous namespace\)::|cmh::|void ", "", name)
    short = re.sub(r"\(.*", "", short)
    if len(sys.argv) > 1 and not re.search(sys.argv[1], short):
        continue
    print("%-62s %s" % (short, "  ".join("%s x%d" % kv for kv in sorted(per[raw].items()))))
