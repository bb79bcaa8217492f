"""Multi-GPU correctness ON HARDWARE (needs >= 2 GPUs: `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`;
skipped on a 1-GPU box).  One process per GPU under torchrun / NCCL (tests/_multi_gpu_worker.py): sharded records
concatenate bit-exactly into the single-GPU records, the NCCL moment all-reduce gives the single-GPU RMS spot radius,
and the all-reduced parameter gradients of the sharded autograd step equal the single-GPU backward pass."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.timeout(900)
def test_sharded_trace_and_gradient_step_equal_single_gpu():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "_multi_gpu_worker.py")]
    res = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=850)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    line = [ln for ln in res.stdout.splitlines() if ln.startswith("MULTI_GPU_RESULT ")][-1]
    out = json.loads(line[len("MULTI_GPU_RESULT "):])
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "multi_gpu_correctness.json"), "w") as f:
        json.dump(out, f, indent=1)
    assert out["world"] == world
    for k in ("dgauss_c2_float32", "dgauss_c2_float64", "hubble_c4_float64"):
        assert out[k]["bit_identical"], k
        assert out[k]["rms_rel_err"] <= 1e-9, (k, out[k])
    assert out["c5_polarized_f64"]["intensity_bit_identical"] and out["c5_polarized_f64"]["P_bit_identical"]
    assert out["c3_sharded_gradient"]["loss_rel_err"] <= 1e-12
    assert out["c3_sharded_gradient"]["grad_max_abs_err_over_scale"] <= 1e-10
