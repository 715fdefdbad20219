"""CPU oracle for the deplex plane-extraction hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs may
import this package.  Nothing under deplex_b200/ does.
"""
from .oracle import (  # noqa: F401
    OracleConfig, OracleError, build, depth_to_cloud, eig3, load_ini, process, process_batch, ref_dsyev_path,
    set_sum_variant,
)
