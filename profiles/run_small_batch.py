"""Workload for the small-batch ncu captures: the 20x256 net evaluated at n = 1 and n = 16 through the packed entry
(encoder -> conv3x3_sb tower -> last layer on conv3x3_v2 with fused heads -> head FC).  Each call is one CUDA graph
launch of 44 kernels; 4 calls per batch size."""
import sys
import tempfile

sys.path.insert(0, ".")
from unrealgo_b200 import DlTFNetworkEvaluator, synth, tfformat  # noqa: E402

d = tempfile.mkdtemp(prefix="ug_prof_")
meta, prefix = tfformat.write_synthetic_checkpoint(d, 20, 256, seed=20260119)
ev = DlTFNetworkEvaluator(meta, max_batch=32)
assert ev.UpdateCheckPoint(prefix)
packed = synth.synthetic_packed_positions(16, seed=3)
for n in (1, 16):
    for _ in range(4):
        ev.evaluate_packed(packed[:n])
ev.close()
