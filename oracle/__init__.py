"""oracle/ — TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference hot path.

PARITY UNPINNED: the reference (ossamaAhmed/blackbox_mpc @ 68c9e63) ships no tests, golden
vectors or fixtures, and its arithmetic lives in TensorFlow 2.0.0 which cannot be installed in
this image (no cp312 wheel, no network).  This package restates the reference's algorithm op by
op (every function cites the reference file:line it follows) in torch-CPU; it is pinned only
against closed forms and self-generated vectors (tests/golden/, script committed).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package, and only as the checker or the timed CPU baseline.  The product package
(blackbox_mpc_b200) never imports it.
"""
from .reference_port import *  # noqa: F401,F403
