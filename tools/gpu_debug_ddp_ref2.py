#!/usr/bin/env python
"""run-to-run difference of the two-shard single-process reference (tests/test_ddp_gpu.py) under the current environment"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import test_ddp_gpu as T
w = int(sys.argv[1]) if len(sys.argv) > 1 else 2
a = T._single_process_reference('krn', w)
b = T._single_process_reference('krn', w)
d = (a.double() - b.double()).abs()
print('shards %d: max %.3e median %.3e frac>1e-4 %.3e' % (w, float(d.max()), float(d.median()), float((d > 1e-4).double().mean())), {k: v for k, v in os.environ.items() if k.startswith('B200SP_')})
