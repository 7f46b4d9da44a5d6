"""Smoke test of the tcgen05 k-NN filter on small inputs (GPU box): estimate error + parity."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from flake16_framework_b200 import ops
rs = np.random.RandomState(0)
for n, nq, d in ((1000, 300, 16), (5000, 5000, 7), (20000, 20000, 16)):
    A = torch.from_numpy(rs.randn(n, d)).cuda()
    Q = torch.from_numpy(rs.randn(nq, d)).cuda()
    print("n=%d nq=%d d=%d" % (n, nq, d), flush=True)
    print("  probe mma.sync  err = %.3e" % ops.knn_tc_probe(A, Q), flush=True)
    print("  probe tcgen05   err = %.3e" % ops.knn_tc_probe(A, Q, umma=True), flush=True)
    order = np.arange(d, dtype=np.int32)
    ref = ops.knn(A, Q, 4, (order, 0)).cpu().numpy()
    got = ops.knn(A, Q, 4, (order, 4)).cpu().numpy()
    print("  parity A!=Q rows differing: %d of %d" % ((got != ref).any(axis=1).sum(), nq), flush=True)
    ref = ops.knn(A, A, 4, (order, 0)).cpu().numpy()
    got = ops.knn(A, A, 4, (order, 4)).cpu().numpy()
    print("  parity self  rows differing: %d of %d" % ((got != ref).any(axis=1).sum(), n), flush=True)
print("UMMA_PROBE_DONE")
