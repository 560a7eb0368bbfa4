"""Small workload meant for compute-sanitizer (memcheck / initcheck / synccheck): both tower kernels, the head kernels
with and without a legal mask (dense and compact), the leaf queue with a batch that is padded to the next multiple of
8 (rows past n must be defined data), the reference-facing Evaluate() path.
    compute-sanitizer --tool initcheck python tests/gpu_sanitizer_probe.py
(compute-sanitizer is closed on this project's GPU pool — it left GPUs needing a reset — so here the script only runs
plain, as a self-checking smoke of the same paths.)"""
import sys
import tempfile

import numpy as np

sys.path.insert(0, ".")
from unrealgo_b200 import DlTFNetworkEvaluator, LeafQueue, synth, tfformat  # noqa: E402

d = tempfile.mkdtemp(prefix="ug_san_")
meta, prefix = tfformat.write_synthetic_checkpoint(d, 1, 64, seed=3)
ev = DlTFNetworkEvaluator(meta, max_batch=48)
assert ev.UpdateCheckPoint(prefix)
packed = synth.synthetic_packed_positions(40, seed=1)
planes = synth.packed_to_planes(packed)
legal = np.full((40, 12), 0xFFFFFFFF, np.uint32)
legal[:, 11] = 0x3FF
legal[::2, 3] = 0x0F0F0F0F
p40, v40 = ev.evaluate_packed(packed)            # conv3x3_v2
p3, v3 = ev.evaluate_packed(packed[:3])           # conv3x3_sb
assert np.array_equal(p3, p40[:3])
pl, _ = ev.evaluate_packed(packed[:5], legal[:5])
pol, val = np.zeros((7, 362)), np.zeros(7)
assert ev.Evaluate(planes[:7], pol, val, 7)
q = LeafQueue(ev, max_batch=16, timeout_us=200)
q.set_compact_priors(True)
t = [q.submit(packed[i], legal[i]) for i in range(5)]   # 5 leaves -> launched as 8
for i, tk in enumerate(t):
    r, _ = q.wait(tk)
    k = int(sum(bin(int(w)).count("1") for w in legal[i]))
    assert abs(float(r[:k].sum()) - 1) < 1e-4
q.close()
ev.close()
print("sanitizer probe ok")
