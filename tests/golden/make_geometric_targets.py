#!/usr/bin/env python
"""Extracts the Monte-Carlo target coefficients v_mc_225_40(1:20, :) of the reference's own test
tests/test_geometric_coeffs/test_geometric_coeffs.F90:128-221 into tests/golden/geometric_coeffs_mc_225_40.txt
(20 rows x 9 values; the reference tree is only needed to regenerate the fixture).
    python tests/golden/make_geometric_targets.py /root/reference"""
import os
import re
import sys

import numpy as np

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
src = open(os.path.join(ref, "tests/test_geometric_coeffs/test_geometric_coeffs.F90")).read()
rows = {}
for m in re.finditer(r"v_mc_225_40\((\d+), :\) = \[real\(ireals\) :: &(.*?)\]", src, flags=re.S):
    vals = [float(x) for x in re.findall(r"[-+]?\d+\.\d+(?:[eE][-+]?\d+)?", m.group(2))]
    assert len(vals) == 9, (m.group(1), vals)
    rows[int(m.group(1))] = vals
assert sorted(rows) == list(range(1, 21))
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "geometric_coeffs_mc_225_40.txt")
np.savetxt(out, np.array([rows[i] for i in range(1, 21)]), fmt="%.9f",
           header="v_mc_225_40(i, 1:9), i = 1..20: boxmc dir2dir of the distorted unit box, phi = 225, theta = 40 "
                  "(tests/test_geometric_coeffs/test_geometric_coeffs.F90:128-221)")
print(out)
