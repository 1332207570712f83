"""Set encoder and FiLM generator behind the reference's interfaces.

Mirrors reference ``model/set_encoders.py`` (``SetEncoder``, ``NullSetEncoder``) and ``model/feature_adapters.py``
(``FilmParameterGenerator``, ``NullGenerator``): same class names, constructor arguments, methods and state-dict keys
(``encoder.layer{i}.{0,1}.*``; ``generators.{i}.block.{0,1,3}.*``, ``regularizers.{i}``). The arithmetic is native:
the set encoder runs through the engine's layer plan (im2col + tcgen05 GEMM + max-pool), the generator is one launch."""
import ctypes as C

import numpy as np
import torch
import torch.nn as nn

from . import lib as L
from .feature_extractors import FeatureExtractor, _Node


class NullSetEncoder(nn.Module):
    """set_encoders.py:122-134."""

    def forward(self, x):
        return None

    def count_macs(self, *inputs):
        return 0

    def aggregate(self, x, aggregation='mean'):
        return None

    @property
    def output_size(self):
        return None


class NullGenerator(nn.Module):
    """feature_adapters.py:80-95."""

    def forward(self, x):
        return {}

    def count_macs(self, *inputs):
        return 0

    def regularization_term(self):
        return 0

    def as_blob(self, film_dict):
        return None


class SetEncoder(FeatureExtractor):
    """set_encoders.py:34-79: frames -> 64-d per-frame embeddings; ``aggregate`` means over all support frames."""

    def __init__(self):
        super().__init__('set_encoder')

    def forward(self, x):
        if x.dim() == 5:
            x = x.flatten(end_dim=1)
        if torch.is_grad_enabled() and self.train_graph:
            # meta-training (single-step-learner.py:196-243): the embeddings carry a graph back to every parameter
            from .training import SetEncoderFn
            params = tuple(p for p in self.parameters() if p.requires_grad)
            if params:
                return SetEncoderFn.apply(self, x, *params)
        return super().forward(x)
    # count_macs: inherited from FeatureExtractor (the set encoder runs through the same engine)

    train_graph = False      # set by the recogniser: True while meta-training (not test mode, autograd on)

    def aggregate(self, x, aggregation='mean'):
        if not isinstance(x, torch.Tensor):
            x = torch.cat(x, dim=0)
        if aggregation == 'mean' and x.requires_grad and torch.is_grad_enabled():
            return x.mean(dim=0, keepdim=True)      # meta-training: autograd routes the mean (set_encoders.py:68-71)
        if aggregation == 'mean':
            out = torch.empty(1, x.shape[1], dtype=torch.float32, device=x.device)
            x = x.contiguous()
            L.check(L.load().orbit_pool_clips(L.ptr(x), 1, x.shape[0], x.shape[1], L.ptr(out), L.stream_ptr(x.device)),
                    "orbit_pool_clips")
            L.count_launches(1)
            return out
        if aggregation == 'none':
            return x
        raise ValueError(f'Aggregation method {aggregation} not valid!')



_TABLE_DTYPE = np.dtype([(k, '<i8') for k in ('w1', 'b1', 'ln_w', 'ln_b', 'w2', 'b2', 'reg', 'init', 'out')] +
                        [('size', '<i4'), ('is_weight', '<i4')])


class FilmParameterGenerator(nn.Module):
    """feature_adapters.py:36-78. ``forward(z)`` returns the reference's ``{name: tensor}`` dict; the tensors are
    views of ONE film blob (sorted-name order, feature_adapters.py:43-44) that the extractor engine folds directly."""

    def count_macs(self, *inputs):
        """Dense-layer MACs of every DenseBlock (mlps.py:52-71: pooled->hidden, hidden->size) plus the regulariser
        scale (one multiply per generated value, feature_adapters.py:62-66)."""
        return sum(self.hidden * self.hidden + self.hidden * r['size'] + r['size'] for r in self._rows)

    def __init__(self, film_parameter_sizes, initial_film_parameters, pooled_size, hidden_size):
        super().__init__()
        if pooled_size != hidden_size:
            raise NotImplementedError("the generator kernel assumes pooled_size == hidden_size (reference: 64 == 64)")
        assert _TABLE_DTYPE.itemsize == L.load().orbit_film_table_entry_bytes()
        self.film_parameter_names = sorted(initial_film_parameters.keys())
        self.film_parameter_sizes = dict(film_parameter_sizes)
        self.hidden = hidden_size
        self.l2_term = 0.0
        self._generation = 0
        self._initial_key = None
        # flat parameter blob: every tensor 16-byte aligned; nn.Parameters are views (as in FeatureExtractor)
        layout, off = [], 0

        def add(name, shape):
            nonlocal off
            n = int(np.prod(shape))
            layout.append((name, shape, off, n))
            off += (n + 3) // 4 * 4
            return layout[-1][2]

        self._rows = []
        for i, name in enumerate(self.film_parameter_names):
            size = film_parameter_sizes[name]
            pre = f'generators.{i}.block.'
            row = dict(w1=add(pre + '0.weight', (hidden_size, pooled_size)), b1=add(pre + '0.bias', (hidden_size,)),
                       ln_w=add(pre + '1.weight', (hidden_size,)), ln_b=add(pre + '1.bias', (hidden_size,)),
                       w2=add(pre + '3.weight', (size, hidden_size)), b2=add(pre + '3.bias', (size,)),
                       reg=add(f'regularizers.{i}', (size,)), init=add(f'_initial.{i}', (size,)),
                       size=size, is_weight=1 if 'weight' in name else 0)
            self._rows.append(row)
        self._layout = layout
        self._blob = torch.zeros(off, dtype=torch.float32)
        self._views = {}
        g = torch.Generator().manual_seed(0)
        for name, shape, o, n in layout:
            view = self._blob[o:o + n].view(shape)
            if name.startswith('_initial.'):
                continue
            parent, parts = self, name.split('.')
            for p in parts[:-1]:
                if not hasattr(parent, p):
                    parent.add_module(p, _Node())
                parent = getattr(parent, p)
            if name.startswith('regularizers.'):
                view.copy_(torch.randn(shape, generator=g) * 0.001)           # feature_adapters.py:50
            elif name.endswith('1.weight'):
                view.fill_(1.0)
            elif name.endswith('.weight'):
                view.copy_((torch.rand(shape, generator=g) * 2 - 1) * shape[1] ** -0.5)   # nn.Linear default range
            elif name.endswith(('0.bias', '3.bias')):
                view.copy_((torch.rand(shape, generator=g) * 2 - 1) * hidden_size ** -0.5)
            parent.register_parameter(parts[-1], nn.Parameter(view, requires_grad=True))
        self.initial_film_parameters = initial_film_parameters   # plain dict, NOT in state_dict (as the reference)
        self._table_dev = None
        self._film_blob = None
        self._out_offsets = {}
        o = 0
        for name in self.film_parameter_names:
            self._out_offsets[name] = o
            o += film_parameter_sizes[name]
        self._film_floats = o
        self._max_size = max(film_parameter_sizes.values())

    def _rebind(self):
        params = dict(self.named_parameters())
        for name, shape, o, n in self._layout:
            if name in params:
                params[name].data = self._blob[o:o + n].view(shape)

    def _apply(self, fn, recurse=True):
        # moves the blob (and, like the reference's custom _apply, the plain initial_film_parameters dict)
        self._blob = fn(self._blob)
        self._rebind()
        self.initial_film_parameters = {k: fn(v) for k, v in self.initial_film_parameters.items()}
        self._table_dev = self._film_blob = None
        return self

    def regularization_term(self):
        return self.l2_term

    def as_blob(self, film_dict):
        """The film blob behind a dict returned by forward() (None for an empty dict)."""
        return self._film_blob if film_dict else None

    def _param_slots(self):
        params = dict(self.named_parameters())
        return {id(params[name]): (o, n, shape) for name, shape, o, n in self._layout if name in params}

    def forward(self, x):
        if x.requires_grad and torch.is_grad_enabled():
            # meta-training: film blob with a graph back to the generator's parameters and the task embedding
            from .training import FilmGeneratorFn
            params = tuple(p for p in self.parameters() if p.requires_grad)
            film = FilmGeneratorFn.apply(self, x, *params)
            film._orbit_generation = (id(self), self._generation)
            self._film_blob = film
            regs = [p for n, p in self.named_parameters() if n.startswith('regularizers.')]
            self.l2_term = sum((r ** 2).sum() for r in regs)        # feature_adapters.py:76, differentiable
        else:
            film = self._generate(x)
            self._film_blob = film
            # l2 term of the regularisers (feature_adapters.py:76): depends on parameters only, not on the episode
            regs = [p for n, p in self.named_parameters() if n.startswith('regularizers.')]
            self.l2_term = sum((r.detach() ** 2).sum() for r in regs)
        return {name: film[self._out_offsets[name]:self._out_offsets[name] + self.film_parameter_sizes[name]]
                for name in self.film_parameter_names}

    def _generate(self, x):
        lib = L.load()
        L.require_cuda(x, "task embedding")
        dev = x.device
        z = x.reshape(-1).contiguous().float()
        assert z.numel() == self.hidden
        # gamma0 / beta0 snapshot lives in the blob next to the generator weights (re-copied only when the dict changed)
        init_key = (self._blob.data_ptr(), tuple((id(v), v._version) for v in self.initial_film_parameters.values()))
        if init_key != self._initial_key:
            for i, name in enumerate(self.film_parameter_names):
                row = self._rows[i]
                self._blob[row['init']:row['init'] + row['size']].copy_(self.initial_film_parameters[name])
            self._initial_key = init_key
        if self._table_dev is None:
            tab = np.zeros(len(self._rows), dtype=_TABLE_DTYPE)
            for i, (row, name) in enumerate(zip(self._rows, self.film_parameter_names)):
                for k in ('w1', 'b1', 'ln_w', 'ln_b', 'w2', 'b2', 'reg', 'init', 'size', 'is_weight'):
                    tab[i][k] = row[k]
                tab[i]['out'] = self._out_offsets[name]
            self._table_dev = torch.from_numpy(tab.view(np.uint8).copy()).to(dev)
        film = torch.empty(self._film_floats, dtype=torch.float32, device=dev)
        L.check(lib.orbit_film_generate(L.ptr(self._blob), L.ptr(self._table_dev), len(self._rows), self._max_size, L.ptr(z),
                                        self.hidden, L.ptr(film), L.stream_ptr(dev)), "orbit_film_generate")
        L.count_launches(1)
        self._generation += 1
        film._orbit_generation = (id(self), self._generation)   # the kernel wrote through a raw pointer: no version bump
        return film
