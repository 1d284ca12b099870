"""Host-side runtime helpers around the scoring path (sharding of clips / images over ranks)."""
from .sharding import ShardPlan, bind_to_gpu_numa_node, gather_scores, shard_range  # noqa: F401
