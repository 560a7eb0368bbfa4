"""TEST INFRASTRUCTURE — a stand-in for tools/tf_dump_reference.py where TensorFlow does not exist.

Writes the same .npz layout from the float64 numpy oracle (oracle/net_oracle_np.py) instead of TensorFlow, marked
producer = "oracle-standin", so that tools/compare_with_tf.py can be exercised end to end on the GPU box
(tests/test_gpu_tf_harness.py).  A stand-in dump pins nothing about TensorFlow; it proves the harness runs."""
import json

import numpy as np


def write(out_path, meta, prefix, planes, packed=None, value_transform=0, input_name="feature_input", outputs=None):
    from oracle.net_oracle_np import NumpyGraphOracle
    outputs = outputs or ["resnet/tower_0/policy_head/policy_predict", "resnet/tower_0/value_head/reward_predict"]
    orc = NumpyGraphOracle(meta, prefix, input_name, outputs)
    pol, val = orc.evaluate_reference(planes, 0)
    v64 = -val if value_transform == 1 else (1 - val if value_transform == 2 else val)
    info = {"producer": "oracle-standin", "tensorflow_version": None, "input": input_name, "outputs": list(outputs),
            "input_dtype": "bool", "value_transform": value_transform, "positions": int(planes.shape[0])}
    save = {"planes": np.ascontiguousarray(planes, np.int8), "policy": pol.astype(np.float32), "value": val.astype(np.float32),
            "value_transformed": np.asarray(v64, np.float64), "info": np.array(json.dumps(info))}
    if packed is not None:
        save["packed"] = packed
    np.savez_compressed(out_path, **save)
    return out_path
