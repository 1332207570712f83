"""Sharded evaluation: episodes are independent, so they are dealt round-robin to the
ranks and the only collective of a run is ONE all-reduce(sum) of a handful of metric accumulators.

Metric definitions follow reference utils/eval_metrics.py: frame accuracy of a video = mean(argmax == label)
(:27-36), the reported statistic is the mean over videos with a 95% confidence interval 1.96*std/sqrt(n)
(:24-25, :213-219). Sums are kept as float64 / int64 so that the result is identical for 1, 2, 4 or 8 ranks."""
import math

import torch


def shard_episodes(num_episodes: int, rank: int, world_size: int):
    """Episode e runs on rank e mod world_size (episodes are seeded by index, never by rank)."""
    return range(rank, num_episodes, world_size)


class ShardedFrameAccuracy:
    """Running sums [correct frames, frames, sum acc_v, sum acc_v^2, videos]; ``append_video`` never syncs."""

    def __init__(self, device):
        self.counts = torch.zeros(2, dtype=torch.int64, device=device)
        self.sums = torch.zeros(3, dtype=torch.float64, device=device)

    def append_video(self, predictions: torch.Tensor, label):
        """``predictions``: [frames] arg-max class indices (or [frames, C] logits) of ONE video."""
        if predictions.dim() == 2:
            predictions = predictions.argmax(dim=-1)
        correct = (predictions.long() == label).sum()
        n = predictions.numel()
        acc = correct.double() / n
        self.counts += torch.stack((correct, torch.tensor(n, device=correct.device)))
        self.sums += torch.stack((acc, acc * acc, torch.ones((), dtype=torch.float64, device=acc.device)))

    def reduce(self, group=None):
        """All-reduce over the ranks (NCCL on GPUs, gloo on CPU) and return the run's statistics."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.counts, group=group)
            dist.all_reduce(self.sums, group=group)
        correct, frames = (int(v) for v in self.counts.tolist())
        s1, s2, n = self.sums.tolist()
        mean = s1 / n if n else 0.0
        var = max(s2 / n - mean * mean, 0.0) if n else 0.0        # np.std (population), eval_metrics.py:24-25
        return {"correct_frames": correct, "frames": frames, "videos": int(n),
                "frame_acc_mean_over_videos": mean, "frame_acc_ci95": 1.96 * math.sqrt(var) / math.sqrt(n) if n else 0.0}
