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

    def append_videos(self, predictions: torch.Tensor, labels: torch.Tensor):
        """``predictions`` [videos, frames] arg-max class indices of equally long videos, ``labels`` [videos]: the same
        sums as calling append_video per row, in four device operations and without a host sync."""
        hit = predictions.long() == labels.long().unsqueeze(1)
        correct = hit.sum(dim=1)
        acc = correct.double() / predictions.shape[1]
        self.counts += torch.stack((correct.sum(), torch.tensor(predictions.numel(), device=correct.device)))
        self.sums += torch.stack((acc.sum(), (acc * acc).sum(), torch.tensor(float(len(acc)), dtype=torch.float64, device=acc.device)))

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


# ------------------------------------------------------------------------------------------------------------------
# Device-side evaluators with the reference's interface (utils/eval_metrics.py:14-352)
# ------------------------------------------------------------------------------------------------------------------
import json  # noqa: E402
from pathlib import Path  # noqa: E402

import numpy as np  # noqa: E402

from . import lib as L  # noqa: E402
from .ops_counter import OpsCounter, clever_format  # noqa: E402

_STATS = ('frame_acc', 'frames_to_recognition', 'video_acc')


def _confidence_interval(scores):
    """eval_metrics.py:24-25."""
    return (1.96 * np.std(scores)) / np.sqrt(len(scores))


def _group_score(stat, rows):
    """Score of the concatenation of the videos in ``rows`` ([k,5] int64: correct, frames, first, mode, label), i.e.
    what the reference obtains by applying its stat function to the flattened frame list (eval_metrics.py:177-197)."""
    correct, frames, first = rows[:, 0], rows[:, 1], rows[:, 2]
    total = int(frames.sum())
    if stat == 'frame_acc':                      # np.mean(correct flags): an exact integer sum / count in float64
        return float(np.float64(int(correct.sum())) / np.float64(total))
    if stat == 'frames_to_recognition':          # first correct frame of the concatenation / its length; 1.0 if none
        hit = np.nonzero(first < frames)[0]
        if len(hit) == 0:
            return 1.0
        v = int(hit[0])
        return float(np.float64(int(frames[:v].sum()) + int(first[v])) / np.float64(total))
    if len(rows) == 1:                           # video_acc (eval_metrics.py:38-46)
        return 1.0 if rows[0, 3] == rows[0, 4] else 0.0
    # the reference compares one modal prediction with an ARRAY of labels here, which numpy refuses to reduce to a bool
    raise ValueError("video_acc is only defined for a single video (the truth value of an array with more than one "
                     "element is ambiguous)")


class _DeviceVideoStats:
    """Growable device table of per-video integer statistics, filled by ``orbit_video_stats``; never syncs on append."""

    def __init__(self):
        self.n = 0
        self.stats = None       # [cap, 4] int32 on device
        self.labels = None      # [cap] int32 on device
        self.preds = []         # per video: [frames] int32 arg-max tensor (kept on device until ``save``)

    def _reserve(self, device):
        if self.stats is None or self.stats.device != device:
            self.stats = torch.zeros(256, 4, dtype=torch.int32, device=device)
            self.labels = torch.zeros(256, dtype=torch.int32, device=device)
            self._offsets = {}
        if self.n == self.stats.shape[0]:
            self.stats = torch.cat((self.stats, torch.zeros_like(self.stats)))
            self.labels = torch.cat((self.labels, torch.zeros_like(self.labels)))

    def append(self, frame_logits, video_label, frame_index=None):
        L.require_cuda(frame_logits, "frame_logits")
        dev = frame_logits.device
        self._reserve(dev)
        logits = frame_logits.detach().contiguous().float()
        n_scored = int(frame_index.numel()) if frame_index is not None else logits.shape[0]
        key = (n_scored, dev)
        if key not in self._offsets:
            self._offsets[key] = torch.tensor([0, n_scored], dtype=torch.int32).to(dev, non_blocking=True)
        slot = self.n
        label = torch.as_tensor(video_label).reshape(-1)[:1].to(device=dev, dtype=torch.int32, non_blocking=True)
        self.labels[slot:slot + 1].copy_(label)
        pred = torch.empty(n_scored, dtype=torch.int32, device=dev)
        L.check(L.load().orbit_video_stats(L.ptr(logits), None, logits.shape[1], L.ptr(frame_index), L.ptr(self._offsets[key]),
                                           L.ptr(self.labels[slot:slot + 1]), 1, L.ptr(self.stats[slot:slot + 1]), L.ptr(pred),
                                           L.stream_ptr(dev)), "orbit_video_stats")
        L.count_launches(1)
        self.preds.append(pred)
        self.n += 1
        return slot

    def table(self):
        """[n, 5] int64 on the host: correct, frames, first correct, modal prediction, label. ONE device->host read."""
        if self.n == 0:
            return np.zeros((0, 5), dtype=np.int64)
        both = torch.cat((self.stats[:self.n], self.labels[:self.n, None]), dim=1)
        return both.cpu().numpy().astype(np.int64)


class Evaluator:
    """eval_metrics.py:14-68. The stat functions keep the reference's (label, probs) host signature for direct use."""

    def __init__(self, stats_to_compute):
        for stat in stats_to_compute:
            if stat not in _STATS:
                raise KeyError(stat)
        self.stats_to_compute = stats_to_compute
        self.stat_fns = {'frame_acc': self.get_frame_accuracy, 'frames_to_recognition': self.get_frames_to_recognition,
                         'video_acc': self.get_video_accuracy}

    def get_confidence_interval(self, scores):
        return _confidence_interval(scores)

    @staticmethod
    def _rows(label, probs):
        pred = np.argmax(np.asarray(probs), axis=-1).reshape(-1)
        ok = np.nonzero(pred == np.asarray(label).reshape(-1) if np.ndim(label) else pred == label)[0]
        first = int(ok[0]) if len(ok) else len(pred)
        lab = int(np.asarray(label).reshape(-1)[0])
        return np.array([[len(ok), len(pred), first, int(np.bincount(pred).argmax()), lab]], dtype=np.int64)

    def get_frame_accuracy(self, label, probs):
        return _group_score('frame_acc', self._rows(label, probs))

    def get_frames_to_recognition(self, label, probs):
        return _group_score('frames_to_recognition', self._rows(label, probs))

    def get_video_accuracy(self, label, probs):
        return _group_score('video_acc', self._rows(label, probs))

    def get_video_prediction(self, probs):
        return int(np.bincount(np.argmax(np.asarray(probs), axis=-1)).argmax())


class TrainEvaluator(Evaluator):
    """eval_metrics.py:70-98: statistics of one batch of target clips, every clip with its own label. The arg-max and
    the comparison run on the device (one launch, each clip = a one-frame video); the per-batch read of two integers is
    the sync the reference also has (it moves the whole probability matrix)."""

    def __init__(self, stats_to_compute):
        super().__init__(stats_to_compute)
        self.reset()

    def reset(self):
        self.current_stats = {stat: 0.0 for stat in self.stats_to_compute}
        self.running_stats = {stat: [] for stat in self.stats_to_compute}

    def update_stats(self, logits, labels):
        L.require_cuda(logits, "logits")
        dev = logits.device
        n = logits.shape[0]
        lg = logits.detach().contiguous().float()
        offsets = torch.arange(n + 1, dtype=torch.int32, device=dev)
        lab = labels.to(device=dev, dtype=torch.int32)
        stats = torch.empty(n, 4, dtype=torch.int32, device=dev)
        L.check(L.load().orbit_video_stats(L.ptr(lg), None, lg.shape[1], None, L.ptr(offsets), L.ptr(lab), n, L.ptr(stats), None,
                                           L.stream_ptr(dev)), "orbit_video_stats")
        L.count_launches(1)
        t = stats.cpu().numpy().astype(np.int64)
        for stat in self.stats_to_compute:
            if stat == 'frame_acc':
                score = float(np.float64(int(t[:, 0].sum())) / np.float64(n))
            elif stat == 'frames_to_recognition':
                hit = np.nonzero(t[:, 0] > 0)[0]
                score = float(hit[0] / n) if len(hit) else 1.0
            else:
                raise ValueError("video_acc is only defined for a single video")
            self.current_stats[stat] = score
            self.running_stats[stat].append(score)

    def get_current_stats(self):
        return self.current_stats

    def get_mean_stats(self):
        return {stat: [np.mean(s), self.get_confidence_interval(s)] for stat, s in self.running_stats.items()}


class TestEvaluator(Evaluator):
    """eval_metrics.py:100-331. ``append_video`` launches one kernel and returns without touching the host; the
    statistics are formed from ONE read of the integer table in ``get_mean_stats``."""
    __test__ = False   # not a pytest class

    def __init__(self, stats_to_compute, save_dir=None, with_ops_counter=False, count_backwards=False):
        super().__init__(stats_to_compute)
        if save_dir:
            self.save_dir = save_dir
        self.ops_counter = OpsCounter(count_backward=count_backwards) if with_ops_counter else None
        self.reset()

    def reset(self):
        self.current_user = 0
        self.current_task = 0
        self._table = _DeviceVideoStats()
        self.all_video_slots = [[[]]]        # [user][task] -> slots in the device table
        self.all_frame_paths = [[[]]]
        self.all_users = []
        self.all_object_lists = [[[]]]
        self.all_personalise_times = [[[]]]
        self.all_inference_times = [[[]]]
        if self.ops_counter:
            self.macs_counter = [[[]]]
            self.params_counter = [[[]]]

    def append_video(self, frame_logits, video_label, frame_paths=None):
        """eval_metrics.py:260-276. ``frame_paths`` (optional here): duplicates that pad a video to a multiple of the
        clip length are dropped, and frames are scored in sorted-path order, as ``np.unique`` does in the reference."""
        frame_index = None
        if frame_paths is not None:
            frame_paths, unique_idxs = np.unique(np.asarray(frame_paths), return_index=True)
            if len(unique_idxs) != frame_logits.shape[0] or not np.array_equal(unique_idxs, np.arange(len(unique_idxs))):
                frame_index = torch.from_numpy(unique_idxs.astype(np.int32)).to(frame_logits.device, non_blocking=True)
        slot = self._table.append(frame_logits, video_label, frame_index)
        self.all_video_slots[self.current_user][self.current_task].append(slot)
        self.all_frame_paths[self.current_user][self.current_task].append(frame_paths)

    def set_current_user(self, user_id):
        self.all_users.append(user_id)
        assert len(self.all_users) == self.current_user + 1

    def set_task_object_list(self, task_object_list):
        self.all_object_lists[self.current_user][self.current_task] = task_object_list

    def _per_user_lists(self):
        lists = [self.all_video_slots, self.all_frame_paths, self.all_object_lists, self.all_personalise_times,
                 self.all_inference_times]
        if self.ops_counter:
            lists += [self.macs_counter, self.params_counter]
        return lists

    def next_user(self):
        for lst in self._per_user_lists():
            lst.append([[]])
        self.current_task = 0
        self.current_user += 1

    def next_task(self):
        for lst in self._per_user_lists():
            lst[self.current_user].append([])
        self.current_task += 1

    def set_base_params(self, params):
        if self.ops_counter:
            self.ops_counter.set_base_params(params)

    def log_time(self, time, time_type='personalise'):
        if time_type == 'personalise':
            self.all_personalise_times[self.current_user][self.current_task] = time
        elif time_type == 'inference':
            self.all_inference_times[self.current_user][self.current_task] = time
        else:
            raise ValueError(f"time_type must be 'personalise' or 'inference' but got {time_type}")

    def task_complete(self):
        if self.ops_counter:
            self.macs_counter[self.current_user][self.current_task] = self.ops_counter.get_task_macs()
            self.params_counter[self.current_user][self.current_task] = self.ops_counter.get_task_params()
            self.ops_counter.task_complete()

    def check_for_uncounted_modules(self, model):
        if self.ops_counter:
            missing = sorted({type(m).__name__ for m in (model.feature_extractor, model.classifier,
                                                         getattr(model, 'set_encoder', None), getattr(model, 'film_generator', None))
                              if isinstance(m, torch.nn.Module) and any(True for _ in m.parameters())
                              and not hasattr(m, 'count_macs') and not hasattr(type(m), '_count_class_reps')})
            return "MACs are counted analytically by every native module." if not missing else \
                "MACs from these modules will not be counted: " + ", ".join(missing)
        return "TestEvaluator has no ops_counter - cannot check if MACs of all modules will be counted."

    def get_mean_stats(self, current_user=False):
        """eval_metrics.py:155-219: per-user / per-object / per-task / per-video means with 95% confidence intervals."""
        table = self._table.table()
        scores = [{stat: [] for stat in self.stats_to_compute} for _ in range(4)]      # user, object, task, video
        users = [self.current_user] if current_user else range(self.current_user + 1)
        for stat in self.stats_to_compute:
            for user in users:
                user_rows, by_object = [], {}
                for task_slots in self.all_video_slots[user]:
                    rows = table[task_slots] if len(task_slots) else table[:0]
                    for r in rows:
                        scores[3][stat].append(_group_score(stat, r[None]))
                        by_object.setdefault(int(r[4]), []).append(r)
                    scores[2][stat].append(_group_score(stat, rows))
                    user_rows.append(rows)
                for obj_rows in by_object.values():
                    obj = np.stack(obj_rows)
                    scores[1][stat].append(_group_score(stat, obj))
                scores[0][stat].append(_group_score(stat, np.concatenate(user_rows) if user_rows else table[:0]))
        return tuple(self.average_over_scores(s) for s in scores)

    def average_over_scores(self, user_stats):
        return {stat: [np.mean(user_stats[stat]), self.get_confidence_interval(user_stats[stat])]
                for stat in self.stats_to_compute}

    def get_mean_ops_counter_stats(self, current_user=False):
        if not self.ops_counter:
            return "0.00B", "0.00B", "0.00B", ""
        users = [self.current_user] if current_user else range(self.current_user + 1)
        task_macs = [tm for user in users for tm in self.macs_counter[user]]
        task_params = [tp for user in users for tp in self.params_counter[user]]
        mean_ops, std_ops, mean_params = clever_format([np.mean(task_macs), np.std(task_macs), np.mean(task_params)])
        return mean_ops, std_ops, mean_params, self.ops_counter.params_break_down

    def get_mean_times(self, current_user=False):
        """Seconds per task (personalise) and per frame (inference): mean and std over users. The reference formats
        these as strings with utils/logging.py helpers; the raw float seconds are returned here."""
        users = [self.current_user] if current_user else range(self.current_user + 1)
        p = [np.mean(self.all_personalise_times[u]) for u in users]
        i = [np.mean(self.all_inference_times[u]) for u in users]
        return np.mean(p), np.std(p), np.mean(i), np.std(i)

    def save(self):
        """results.json in the reference's layout (eval_metrics.py:111-153): user -> tasks -> videos -> {frame id: prediction}."""
        output = {}
        assert len(self.all_users) == self.current_user + 1
        for user, user_id in enumerate(self.all_users):
            output[user_id] = []
            for task, slots in enumerate(self.all_video_slots[user]):
                task_output = {'task_object_list': self.all_object_lists[user][task], 'task_videos': {}}
                if self.ops_counter:
                    task_output['task_macs_to_personalise'] = int(self.macs_counter[user][task])
                for slot, paths in zip(slots, self.all_frame_paths[user][task]):
                    preds = self._table.preds[slot].cpu().tolist()
                    if paths is None:
                        raise ValueError("save() needs the frame paths of every appended video")
                    assert len(paths) == len(preds)
                    video_id = Path(paths[0]).parts[-2]
                    task_output['task_videos'][video_id] = {int(Path(p).stem.split('-')[-1]): pr for p, pr in zip(paths, preds)}
                output[user_id].append(task_output)
        self.json_results_path = Path(self.save_dir, "results.json")
        self.json_results_path.parent.mkdir(exist_ok=True, parents=True)
        with open(self.json_results_path, 'w') as json_file:
            json.dump(output, json_file)


class ValidationEvaluator(TestEvaluator):
    """eval_metrics.py:333-352."""
    __test__ = False

    def __init__(self, stats_to_compute):
        super().__init__(stats_to_compute)
        self.comparison_stat = self.stats_to_compute[0]
        self.current_best_stats = {stat: [0.0, 0.0] for stat in self.stats_to_compute}

    def is_better(self, stats):
        return bool(stats[self.comparison_stat][0] > self.current_best_stats[self.comparison_stat][0])

    def replace(self, stats):
        self.current_best_stats = stats
