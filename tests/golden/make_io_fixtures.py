"""Write tests/golden/*.sg / *.el with the REFERENCE's own writer (gapbs/writer.h through oracle/ref_shim.cpp), so
the file-format readers can be tested without the reference tree.  Run in the build container:
    python tests/golden/make_io_fixtures.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import binding as B  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
R = B.reference()
assert R is not None
# testing/testGraphs/triangles_3.el, symmetrised (undirected .sg) and as given (directed .sg with inverse)
s = [0, 0, 1, 1, 2, 5, 5, 6, 6, 7, 8]
d = [1, 2, 2, 3, 3, 6, 7, 7, 8, 9, 9]
und = R.from_el(s, d, True)
R.write_file(und, os.path.join(HERE, "triangles_3_undirected.sg"), True)
R.write_file(und, os.path.join(HERE, "triangles_3_undirected.el"), False)
dr = R.from_el(s, d, False)
R.write_file(dr, os.path.join(HERE, "triangles_3_directed.sg"), True)
kron = R.generate(8, 16, False)
R.write_file(kron, os.path.join(HERE, "kronecker_8.sg"), True)
for f in sorted(os.listdir(HERE)):
    if f.endswith((".sg", ".el")):
        print(f, os.path.getsize(os.path.join(HERE, f)))
