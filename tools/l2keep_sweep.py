#!/usr/bin/env python
"""perform_op time against the size of the L2 window the first half leaves behind for the second
(FPB_L2_KEEP_MB), per shape; child process of tools/gather_sweep.py does the timing."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
child = os.path.join(ROOT, "tools", "gather_sweep.py")
shapes = [(10000, 100000, 100000), (500000, 100000, 12500), (500000, 100000, 100000)]
for n, ptot, p in shapes:
    print("== %d x %d (of %d SNPs)" % (n, p, ptot), flush=True)
    for mb in (0, 24, 48, 64, 80, 96, 112):
        e = dict(os.environ, FPB_L2_KEEP_MB=str(mb))
        out = subprocess.run([sys.executable, child, "child", str(n), str(ptot), str(p)], env=e,
                             capture_output=True, text=True)
        print("keep %3d MB  %s" % (mb, out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-400:]),
              flush=True)
