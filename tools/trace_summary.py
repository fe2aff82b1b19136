"""Summarises a DMF_TRACE file (development tool): python tools/trace_summary.py <trace> [first] [last]
Per update u: the window in which moments(u) may run (eligible .. ncc(u) wants to start), when it actually ended, and how long
ncc(u) had to wait for it."""
import sys
import numpy as np
t = np.loadtxt(sys.argv[1], dtype=np.int64, ndmin=2)
a, b = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (0, len(t))
t = t[a:b]
u, adv0, adv1, me, mend, ncc0, ncc1 = t.T
us = lambda x: x / 1e3
print(f"updates {u[0]}..{u[-1]}: per update {us((ncc1[-1] - adv0[0]) / len(t)):.1f} us")
print(f"  advance            {us(adv1 - adv0).mean():7.1f} us")
print(f"  ncc                {us(ncc1 - ncc0).mean():7.1f} us")
print(f"  advance_end -> ncc_begin (wait for moments + launch)   {us(ncc0 - adv1).mean():7.1f} us")
print(f"  moments eligible -> end                                {us(mend - me).mean():7.1f} us")
print(f"  moments_end - advance_end (>0: ncc waited for moments) {us(mend - adv1).mean():7.1f} us   positive in {np.mean(mend > adv1):.0%} of updates")
print(f"  ncc_end(u) -> advance_begin(u+1)                       {us(adv0[1:] - ncc1[:-1]).mean():7.1f} us")
print(f"  moments eligible(u) - ncc_begin(u-1)                   {us(me[1:] - ncc0[:-1]).mean():7.1f} us")
print(f"  moments end(u) - ncc_end(u-1)                          {us(mend[1:] - ncc1[:-1]).mean():7.1f} us")
