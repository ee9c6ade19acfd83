import os, sys, time
sys.path.insert(0, "/root/repo")
os.environ["FPB_IRLM_TRACE"] = "1"
from flashpca_b200.synth import SynthSpec
op = SynthSpec(500000, 100000).create_operator()
for _ in range(2):
    t = time.perf_counter(); r = op.pca_block(20, 1e-6, want_vectors=False); print("block", time.perf_counter() - t, r["npasses"])
