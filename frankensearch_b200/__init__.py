"""frankensearch_b200 — the B200 (sm_100a) semantic-tier hot path of frankensearch behind the
reference's own seams: exact f16 cosine scan + top-k, RRF, two-tier blend, query encoders.

Everything numeric runs in libfsgpu.so (hand-written CUDA, include/fsgpu.h).  There is no CPU
fallback; importing succeeds without a GPU so that host logic can be tested, but every compute
entry point raises `SearchError(SubsystemError)` when no CUDA device is usable.
"""
from ._ffi import SearchError  # noqa: F401
from .types import (ClassifiedHits, FusedHit, QueryClass, RrfConfig, ScoredResult, VectorHit,  # noqa: F401
                    ZeroSignalReason, ZeroSignalState, candidate_count, phase1_budgets, scaled_budget)
from .index import GpuVectorIndex  # noqa: F401
from .filter import BitsetFilter, PredicateFilter  # noqa: F401
from .fusion import (RankChanges, blend_two_tier, blend_two_tier_aligned, compute_rank_changes,  # noqa: F401
                     kendall_tau, rrf_fuse)
from .embed import MiniLmEmbedder, Model2VecEmbedder  # noqa: F401
from .sharded import GpuShardedIndex, ShardedGpuIndex, shard_bounds  # noqa: F401
from .searcher import GpuSyncTwoTierSearcher, SyncSearchOutcome, TwoTierConfig  # noqa: F401
from .two_tier import GpuTwoTierIndex  # noqa: F401
from .pipeline import DeviceLexical, DeviceTwoTierSearcher  # noqa: F401
