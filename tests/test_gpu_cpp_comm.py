"""Runs the C++ test of include/Cabana_B200_Comm.hpp (Cabana::Halo / Distributor / gather /
scatter / migrate templated on the Nccl comm-space tag): one process per GPU, ncclCommInitRank
with an id passed through a file.  world = 1 always runs (one-GPU box: every exchange is the
self block); world = 2 runs when two GPUs are visible."""
import os
import subprocess
import tempfile

import pytest
import torch

pytestmark = pytest.mark.gpu


def _run_world(world):
    from cabana_b200 import build

    exe = build.build_cpp_comm_test()
    with tempfile.TemporaryDirectory() as tmp:
        idf = os.path.join(tmp, "nccl_id")
        env = dict(os.environ)
        env.setdefault("NCCL_DEBUG", "WARN")
        procs = [subprocess.Popen([exe, str(r), str(world), idf], stdout=subprocess.PIPE,
                                  stderr=subprocess.STDOUT, text=True, env=env) for r in range(world)]
        outs = []
        for p in procs:
            try:
                out, _ = p.communicate(timeout=300)
            except subprocess.TimeoutExpired:
                for q in procs:
                    q.kill()
                raise
            outs.append((p.returncode, out))
    everything = "\n".join(f"--- rank {r} rc={rc}\n{out}" for r, (rc, out) in enumerate(outs))
    for r, (rc, out) in enumerate(outs):
        assert rc == 0, f"rank {r} rc={rc}\n{everything}"
        assert f"rank {r}/{world}: ALL CABANA COMM TESTS PASSED" in out, out


def test_cpp_halo_distributor_single_rank():
    assert torch.cuda.is_available()
    _run_world(1)


def test_cpp_halo_distributor_two_ranks():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _run_world(2)
