"""Host-side MAC / parameter accounting with the interface of the reference's ``utils/ops_counter.py:10-99``.

The reference traces modules with ``thop``; here every native module reports its own count analytically
(``module.count_macs(*inputs)``): the feature extractors ask the engine (``orbit_engine_macs``: convolutions, dense
layers, attention matmuls, pooling adds -- normalisation layers and activations are not counted), and the heads use the
closed forms the reference adds by hand (cited at each call site in ``classifier_heads*.py``). Pure host arithmetic: no
kernel is launched for counting.
"""
import torch.nn as nn


def clever_format(values, fmt="%.2f"):
    """Human-readable counts (K/M/G/T suffixes), the formatting the reference takes from thop (ops_counter.py:6,48)."""
    out = []
    for v in values:
        for div, suffix in ((1e12, 'T'), (1e9, 'G'), (1e6, 'M'), (1e3, 'K')):
            if v >= div:
                out.append((fmt % (v / div)) + suffix)
                break
        else:
            out.append((fmt % v) + 'B')
    return out[0] if len(out) == 1 else tuple(out)


def _num_params(module):
    return sum(p.numel() for p in module.parameters()) if isinstance(module, nn.Module) else 0


class OpsCounter:
    """Same attributes and methods as the reference class (ops_counter.py:10-99)."""

    def __init__(self, count_backward=False):
        self.verbose = False
        self.multiplier = 2 if count_backward else 1      # forward + backward (ops_counter.py:13)
        self.task_mac_counter, self.task_params_counter = 0, 0
        self.base_params_counter = 0
        self.params_break_down = ""
        self.personalise_time_per_task = []
        self.inference_time_per_frame = []

    def set_base_params(self, base_model):
        """ops_counter.py:20-48."""
        fe = _num_params(base_model.feature_extractor)
        cl = _num_params(base_model.classifier)
        gen = enc = film = 0
        if base_model.adapt_features:
            if hasattr(base_model, 'film_generator'):
                gen = _num_params(base_model.film_generator)
            if hasattr(base_model, 'set_encoder'):
                enc = _num_params(base_model.set_encoder)
            film = sum(base_model.film_parameter_sizes.values())
        self.base_params_counter = fe + cl + gen + enc + film
        self.params_break_down = "feature extractor: {0:}, classifier: {1:}, film generator: {2:}, set encoder: {3:}, " \
                                 "film params {4:}".format(*clever_format([fe, cl, gen, enc, film]))

    def add_macs(self, num_macs):
        self.task_mac_counter += num_macs

    def add_params(self, num_params):
        self.task_params_counter += num_params

    def compute_macs(self, module, *inputs):
        """ops_counter.py:82-88 with the thop trace replaced by the module's own analytic count. thop's ``params`` of a
        traced module (its parameter count) is added to the task parameters exactly as the reference does."""
        if not hasattr(module, 'count_macs'):
            raise TypeError(f"{type(module).__name__} does not report MACs (no count_macs)")
        self.add_macs(module.count_macs(*inputs) * self.multiplier)
        self.add_params(_num_params(module))

    def task_complete(self):
        self.task_mac_counter = 0
        self.task_params_counter = 0

    def get_task_macs(self):
        return self.task_mac_counter

    def get_task_params(self):
        return self.base_params_counter + self.task_params_counter
