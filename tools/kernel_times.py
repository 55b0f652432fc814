#!/usr/bin/env python3
'''Short IDR solve at grid^3 under `ncu --metrics gpu__time_duration.sum` is the usual use; this script just runs
`products` operator products of the default solver:  python tools/kernel_times.py [grid] [products]'''
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy
from transiflow_b200 import Interface
grid = int(sys.argv[1]) if len(sys.argv) > 1 else 128
prods = int(sys.argv[2]) if len(sys.argv) > 2 else 30
warnings.simplefilter('ignore')
it = Interface({'Reynolds Number': 100, 'Iterative Solver': {'Maximum Iterations': prods, 'Method': 'IDR', 'Schur Complement': 'Scaled Mass'}}, grid, grid, grid)
x = numpy.random.default_rng(0).uniform(-0.01, 0.01, it.n)
jac, f = it.jacobian_rhs(x)
it.solve(jac, -f)
print(it.last_solve)
