"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference evaluator's statistics (utils/eval_metrics.py).

Pinned: ``oracle/make_golden.py`` runs the reference's own ``TestEvaluator`` here (with ``thop``/``timm`` stubbed, they
are only imported for the ops counter) on seeded logits and stores its outputs in ``tests/golden/evaluator.npz``;
``tests/test_oracle_golden.py`` checks this restatement against them.
"""
import numpy as np


def frame_accuracy(label, logits):
    """eval_metrics.py:27-36 (softmax dropped: arg-max is unchanged by it)."""
    return float(np.mean((np.argmax(logits, axis=-1) == label).astype(int)))


def frames_to_recognition(label, logits):
    """eval_metrics.py:48-60."""
    pred = np.argmax(logits, axis=-1)
    hits = np.where(pred == label)[0]
    return float(hits[0] / len(pred)) if len(hits) else 1.0


def video_accuracy(label, logits):
    """eval_metrics.py:38-46,62-68."""
    return 1.0 if np.bincount(np.argmax(logits, axis=-1)).argmax() == label else 0.0


STAT_FNS = {'frame_acc': frame_accuracy, 'frames_to_recognition': frames_to_recognition, 'video_acc': video_accuracy}


def mean_and_ci(scores):
    """eval_metrics.py:24-25,213-219."""
    return [float(np.mean(scores)), float(1.96 * np.std(scores) / np.sqrt(len(scores)))]


def mean_stats(stat, users):
    """eval_metrics.py:155-211. ``users``: list over users of list over tasks of list of (label, logits[F, C]).
    Returns {level: [mean, ci]} for level in user / object / task / video."""
    fn = STAT_FNS[stat]
    per = {'user': [], 'object': [], 'task': [], 'video': []}
    for tasks in users:
        u_labels, u_logits, by_obj = [], [], {}
        for videos in tasks:
            t_labels, t_logits = [], []
            for label, logits in videos:
                per['video'].append(fn(label, logits))
                t_labels.append(np.full(len(logits), label))
                t_logits.append(logits)
                by_obj.setdefault(int(label), []).append(logits)
            per['task'].append(fn(np.concatenate(t_labels), np.concatenate(t_logits)))
            u_labels += t_labels
            u_logits += t_logits
        for obj, chunks in by_obj.items():
            per['object'].append(fn(obj, np.concatenate(chunks)))
        per['user'].append(fn(np.concatenate(u_labels), np.concatenate(u_logits)))
    return {level: mean_and_ci(s) for level, s in per.items()}
