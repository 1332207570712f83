"""Set encoder and FiLM generator behind the reference's interfaces
(reference ``model/set_encoders.py``, ``model/feature_adapters.py``)."""
import torch.nn as nn


class NullSetEncoder(nn.Module):
    """set_encoders.py:122-134."""

    def forward(self, x):
        return None

    def aggregate(self, x, aggregation='mean'):
        return None

    @property
    def output_size(self):
        return None


class NullGenerator(nn.Module):
    """feature_adapters.py:80-95."""

    def forward(self, x):
        return {}

    def regularization_term(self):
        return 0

    def as_blob(self, film_dict):
        return None


class SetEncoder(nn.Module):
    def __init__(self):
        raise NotImplementedError("adapt_features=True (CNAPs set encoder) is not implemented yet")


class FilmParameterGenerator(nn.Module):
    def __init__(self, *a, **k):
        raise NotImplementedError("adapt_features=True (FiLM generator) is not implemented yet")
