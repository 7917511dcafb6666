"""GPU: one full TGANv2 training iteration (B=8, 64x64x16) on the sm_100a kernels against the CPU oracle
on the same weights, inputs and host-RNG stream.  Same thresholds as the CPU emulation test
(tests/test_product_vs_oracle_cpu.py): losses inside the bf16 bar (2e-2); bit-exact real pyramid; gradient
agreement pinned at the level bf16 ReLU-mask noise allows."""
import pytest
import torch

from helpers import golden
from test_product_vs_oracle_cpu import compare, run_product_iteration

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,conditional", [("tganv2_cond_B8.json", True), ("tganv2_uncond_B8.json", False)])
def test_full_iteration_on_b200(name, conditional):
    from txt2vid_b200 import _lib, ops
    ops.PACKS.clear()
    n0 = _lib.lib().t2v_launch_count()
    orc, got = run_product_iteration(conditional, golden(name), "cuda")
    launches = _lib.lib().t2v_launch_count() - n0
    print("kernel launches in one iteration:", launches)
    assert launches > 500
    rep = compare(orc, got, 2e-2, 0.25, 0.97, 6e-2)
    assert rep["gradD"]["l2"] < 5e-2 and rep["gradD"]["cos"] > 0.998, rep
    import json, os
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/iteration_parity_%s.json" % name.split(".")[0], "w") as f:
        json.dump({"launches": int(launches), "report": rep}, f)


def test_config5_128x128x32_iteration_on_b200():
    """BASELINE configs[4]: TGANv2 conditional at 128x128x32 (ConvLSTM plane 2x2, frame sizes 16/32/64/128), one full
    iteration at B = 8 against the oracle on the same weights / inputs / host-RNG stream.  Same bars as the 64x64x16
    test; exercises the kernels on the larger planes (all four pyramid levels take the stride-(2,1,1) stem conv)."""
    from txt2vid_b200 import ops
    ops.PACKS.clear()
    fx = {"config": dict(golden("tganv2_cond_B8.json")["config"])}
    orc, got = run_product_iteration(True, fx, "cuda", size=128, frames=32, frame_sizes=(16, 32, 64, 128))
    rep = compare(orc, got, 2e-2, 0.25, 0.97, 6e-2)
    assert rep["gradD"]["l2"] < 5e-2 and rep["gradD"]["cos"] > 0.998, rep
    import json, os
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/iteration_parity_tganv2_cond_128x128x32_B8.json", "w") as f:
        json.dump({"report": rep}, f)
