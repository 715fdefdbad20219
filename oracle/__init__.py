"""CPU oracle for the deplex plane-extraction hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs may
import this package.  Nothing under deplex_b200/ does.
"""
from .oracle import (  # noqa: F401
    OracleConfig, OracleError, build, depth_to_cloud, eig3, load_ini, process, process_batch, ref_dsyev_path,
    set_sum_variant, set_uniform_int_variant, uniform_selftest, ref_available, ref_full_path, ref_process, ref_cell_stats,
    ref_process_batch,
)
