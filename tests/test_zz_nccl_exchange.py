# -*- coding: utf-8 -*-
"""The interface exchange over NCCL (needs >= 2 GPUs on the box; skipped on a
single-GPU box, where tests/test_distributed.py covers the same code with the
gloo backend).  Named to run last: it is the only multi-process GPU test."""
import os
import socket
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.gpu
def test_interface_exchange_nccl():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
           "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()),
           os.path.join(HERE, "nccl_exchange_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900,
                         env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count("nccl xchg ok") == 5, out.stdout
    assert out.stdout.count("nccl facet ok") == 4, out.stdout
