"""One launch of each tensor-core ranking kernel at BASELINE sizes (for `ncu -k regex:tc_rank_kernel`):
C4-64 hist + rank_topk (10k x 1M, 64 bit, top-1000), then C2 hist + rank_map (5k x 117k, 64 bit, 80 classes)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clip_based_cross_modal_hash_b200 import retrieval as R, synth  # noqa: E402

dev = "cuda"
which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "c4"):
    c = synth.CONFIGS[os.environ.get("C4", "C4-64")]
    qp = R.pack_codes(synth.random_codes(c["Q"], c["K"], 1).to(dev))
    gp = R.pack_codes(synth.random_codes(c["N"], c["K"], 2).to(dev))
    R.topk(qp, gp, c["K"], c["k"])
    torch.cuda.synchronize()
if which in ("all", "c2"):
    c = synth.CONFIGS["C2"]
    qp = R.pack_codes(synth.random_codes(c["Q"], c["K"], 1).to(dev))
    gp = R.pack_codes(synth.random_codes(c["N"], c["K"], 2).to(dev))
    qlp = R.pack_labels(synth.random_labels(c["Q"], c["C"], 3).to(dev))
    glp = R.pack_labels(synth.random_labels(c["N"], c["C"], 4).to(dev))
    print(float(R.map_k(qp, qlp, gp, glp, c["K"], c["C"], None).map))
    torch.cuda.synchronize()
