# -*- coding: utf-8 -*-
"""N > 1 path on the CPU: world_size 2 and 3 with the gloo backend."""
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


@pytest.mark.parametrize("world", [2, 3])
def test_partitioned_assembly_matches_reference(world):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
           "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()),
           os.path.join(HERE, "dist_worker.py")]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    for attempt in range(2):                     # (a rendezvous port can be taken: retry once)
        cmd[cmd.index("--master-port") + 1] = str(_free_port())
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
        if out.returncode == 0:
            break
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    for label, count in (("dist ok", 7), ("xchg ok", 7), ("strip ok", 3), ("facet ok", 21), ("interior facet ok", 3), ("strip facet ok", 4),
                         ("strip exchange ok", 3), ("slab ok", 8), ("slab facet ok", 2)):
        assert out.stdout.count(label) == count, out.stdout
