"""The N > 1 path (z-slab decomposition, halo exchange, dt MAX-reduction).

CPU (-m "not gpu"): world_size-2 and -3 gloo runs of tests/slab_check.py --mode cpu: neighbour bookkeeping of the host Setup
(-mpi / -mpi-s, BC_COPY faces) and the exchange protocol reproduce the ghost planes of the undecomposed block bit for bit.
GPU (-m gpu, needs >= 2 devices; run by hand with `gpurun --gpus 2`): two slabs on two GPUs over NCCL, blocking and overlapped
exchange, against the undecomposed block on one GPU -- bit-exact."""
import os
import socket
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _launch(nproc, args, timeout=600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(HERE, "slab_check.py")] + args
    env = dict(os.environ, OMP_NUM_THREADS="2")
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=timeout, env=env)
    assert r.returncode == 0 and "SLAB_CHECK_OK" in r.stdout, r.stdout[-3000:]
    return r.stdout


@pytest.mark.parametrize("world,case,strong", [(2, "sbi", False), (3, "sbi", False), (2, "jet", True)])
def test_exchange_protocol_gloo(world, case, strong):
    _launch(world, ["--mode", "cpu", "--case", case] + (["--strong"] if strong else []))


@pytest.mark.parametrize("world", [2, 3])
def test_exchange_protocol_periodic_z_gloo(world):
    """Periodic z boundary: the outer faces of the first and last rank exchange with each other; on two ranks both neighbours are the
    same peer and the posting order decides which ghosts a message lands in."""
    _launch(world, ["--mode", "cpu", "--case", "sbi", "--periodic-z"])


@pytest.mark.gpu
@pytest.mark.parametrize("case,strong,weno,pp,alpha", [("sbi", False, 5, 0, "LLF"), ("jet", True, 5, 0, "LLF"), ("sbi", False, 6, 1, "LLF"), ("sbi", False, 7, 0, "ROE"),
                                                       ("sbi", False, 5, 0, "GLF"), ("jet", True, 6, 0, "GLF"), ("sbi-periodic", False, 5, 0, "LLF"),
                                                       ("sbi-visc", False, 5, 0, "LLF"), ("jet-visc", True, 6, 1, "GLF")])
def test_two_slabs_equal_one_block_bitwise(case, strong, weno, pp, alpha):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    per, visc = case.endswith("-periodic"), case.endswith("-visc")
    _launch(2, ["--mode", "gpu", "--case", case.split("-")[0], "--steps", "5", "--weno", str(weno), "--pp", str(pp), "--alpha", alpha] + (["--strong"] if strong else [])
            + (["--periodic-z"] if per else []) + (["--visc"] if visc else []))
