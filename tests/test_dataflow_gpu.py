"""Experimental tile-level hand-over between conv layers: only runs against a library built
with B200_BUILD_DEFINES=B200_DATAFLOW and with B200_SABER_DATAFLOW=1 in the environment; skipped otherwise
(the shipped build does not contain the code path)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _need_dataflow():
    from anakin_b200 import saber_abi
    if not saber_abi.load().b200_dataflow_supported():
        pytest.skip("library built without B200_DATAFLOW or B200_SABER_DATAFLOW != 1")


@pytest.mark.parametrize("batch", [1, 4, 8])
def test_resnet50_int8_dataflow_matches_golden(batch):
    _need_dataflow()
    from anakin_b200 import anakin_bin, api, modelzoo
    gold = np.load(os.path.join(GOLD, "resnet50_golden.npz"))
    g = modelzoo.build("resnet50", batch=batch, precision="int8")
    G = api.Graph.from_bytes(anakin_bin.dumps(g))
    G.ResetBatchSize("input_0", batch)
    G.Optimize()
    net = api.Net(G, "int8")
    net.set_input("input_0", modelzoo.synthetic_input(batch))
    for _ in range(4):          # eager, capture, replays: the counters are re-zeroed every step
        net.prediction()
        net.sync()
        logits, info = net.read_tensor("fc1000")
        c = info["dims"][1]
        logits = (logits[..., :c] if info["layout"] == 9 else logits).reshape(batch, -1)
        n = min(batch, 4)
        np.testing.assert_array_equal(logits[:n], gold["logits_int8"][:n])
