"""Host-side runtime helpers around the scoring path (sharding of clips / images over ranks)."""
from .sharding import (ShardPlan, bind_to_gpu_numa_node, gather_scores, score_clips_sharded,  # noqa: F401
                       shard_range)
