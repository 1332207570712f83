"""Helpers around the hot path; same names and semantics as reference ``data/utils.py``."""
import torch


def get_batch_indices(index, last_element, batch_size):
    """data/utils.py:49-54."""
    batch_start_index = index * batch_size
    batch_end_index = min(batch_start_index + batch_size, last_element)
    return batch_start_index, batch_end_index


def attach_frame_history(frames, history_length):
    """data/utils.py:8-28: clip t = the ``history_length`` frames ending at t, left-padded with
    frame 0. One gather instead of the reference's repeat/roll/stack/slice sequence."""
    if history_length == 1:
        return frames.unsqueeze(1)
    n = frames.shape[0]
    idx = torch.arange(n, device=frames.device)[:, None] + \
        torch.arange(1 - history_length, 1, device=frames.device)[None, :]
    return frames[idx.clamp_min_(0)]


def unpack_task(task_dict, device, context_to_device=True, target_to_device=False):
    """data/utils.py:30-47: only the LABELS move to the device; clips stay where they are."""
    context_labels, target_labels = task_dict['context_labels'], task_dict['target_labels']
    if context_to_device and isinstance(context_labels, torch.Tensor):
        context_labels = context_labels.to(device)
    if target_to_device and isinstance(target_labels, torch.Tensor):
        target_labels = target_labels.to(device)
    return (task_dict['context_clips'], task_dict['context_paths'], context_labels, task_dict['target_clips'],
            task_dict['target_paths'], target_labels, task_dict['object_list'])
