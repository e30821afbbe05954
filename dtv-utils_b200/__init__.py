"""dtv-utils_b200 — B200-native PAPR/CCDF engine, drop-in for the `papr` tool of drmpeg/dtv-utils.

Only the hot path named by BASELINE.json lives here: `csrc/` (hand-written sm_100a kernels and the
C-ABI library, include/papr_b200.h) and `papr.py`, the host-side mirror of the reference's
interface (`papr [-g] <infile>` -> stdout).  Import as `dtv_utils_b200` (see dtv_utils_b200.py at
the repo root: the directory name carries the reference's hyphen).
"""
from .build import build, lib_path, cli_path  # noqa: F401
from .papr import (Engine, MultiEngine, PaprError, PaprResult, PaprStats, analyze_sharded, attach_peer_exchange,  # noqa: F401
                   detach_peer_exchange, format_result, levels, load_library, main, merge_stats)

__all__ = ["build", "lib_path", "cli_path", "Engine", "MultiEngine", "PaprError", "PaprResult", "PaprStats",
           "analyze_sharded", "attach_peer_exchange", "detach_peer_exchange", "format_result", "levels", "load_library", "main", "merge_stats"]
