"""bench.py on the host-layer doubles (tests/host_dry_run.py), launchable by torch.distributed.run: the multi-rank host
path of the bench (sharding, the gradient all-reduce, max-over-ranks timing, the exit path) on `gloo`, without a GPU.
TEST INFRASTRUCTURE ONLY."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import host_dry_run  # noqa: E402

host_dry_run.install()
sys.argv = [os.path.join(ROOT, "bench.py")] + sys.argv[1:]
import bench  # noqa: E402

bench.main()
