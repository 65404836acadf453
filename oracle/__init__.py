"""CPU oracle of the PointNet++ set-abstraction hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under the product package imports this directory; the only
callers are tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.

The reference (0xPabloxx/3d-pointcloud-orientation-estimation) is pure Python/PyTorch, so the oracle
is a numpy / torch-CPU restatement of its algorithm, function by function, each citing the
reference file:line it follows.  It is pinned against the reference itself:
``oracle/make_golden.py`` imports the unmodified reference from /root/reference, checks every oracle
function against it on seeded inputs (indices bit-exact, floating point to 1e-6) and writes the
fixtures under tests/golden/ that the GPU box (which has no /root/reference) replays.  The loss
functions are additionally pinned against the reference's only committed known-answer vectors,
results/multi_peak_vonMises_KL_debug/debug_log.txt.  There is no C component in this oracle.
"""
