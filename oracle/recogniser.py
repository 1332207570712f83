"""personalise()/predict() control flow on CPU (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Restates reference ``model/few_shot_recognisers.py``:
  * FewShotRecogniser._get_features_in_batches (:124-153)  -> OracleRecogniser._features
  * SingleStepFewShotRecogniser.personalise/predict (:313-326, :453-462)
  * MultiStepFewShotRecogniser.personalise/predict (:207-258), utils/optim.py:8-32
All arithmetic is plain PyTorch fp32 on CPU -- the same library path the reference takes with
``--gpu -1`` (single-step-learner.py:65-66).
"""
import math
from collections import OrderedDict

import torch
import torch.nn.functional as F
from torch.func import functional_call

from . import backbones, parts


class OracleRecogniser:
    """One class for both learners. ``classifier``: linear|versa|proto|proto_cosine|mahalanobis
    (few_shot_recognisers.py:72-84)."""

    def __init__(self, feature_extractor_name, adapt_features, classifier, clip_length, batch_size,
                 logit_scale=1.0, seed=1991, calib_input=None):
        if classifier not in ('linear', 'versa', 'proto', 'proto_cosine', 'mahalanobis'):
            raise ValueError(f"Classifier {classifier} not valid.")
        self.name = feature_extractor_name
        self.adapt_features = adapt_features
        self.classifier = classifier
        self.clip_length = clip_length
        self.batch_size = batch_size
        self.logit_scale = logit_scale
        self.extractor = backbones.seeded_init(backbones.build(feature_extractor_name), seed, calib_input)
        self.feat_dim = self.extractor.output_size
        self.film_names = None
        if adapt_features:
            names = parts.film_parameter_names(feature_extractor_name, self.extractor)
            self.film_names = sorted(names)
            ext = dict(self.extractor.named_parameters())
            self.film_initial = {n: ext[n].detach().clone() for n in names}
            self.set_encoder_params = parts.init_set_encoder_params(seed + 1, calib_input)
            self.film_gen_params = parts.init_film_generator_params(
                [self.film_initial[n].numel() for n in self.film_names], seed + 2)
        if classifier == 'versa':
            self.versa_params = parts.init_versa_params(self.feat_dim, seed + 3)
        self.reset()

    # -- state ----------------------------------------------------------------------------
    def reset(self):
        self.film_dict, self.film_l2 = None, 0.0
        self.head = None

    def state_dict(self):
        """Keys as in a reference checkpoint (``feature_extractor.*``, ``set_encoder.*``,
        ``film_generator.*``, ``classifier.*``)."""
        sd = OrderedDict(('feature_extractor.' + k, v.detach().clone())
                         for k, v in self.extractor.state_dict().items())
        if self.adapt_features:
            sd.update(('set_encoder.' + k, v.clone()) for k, v in self.set_encoder_params.items())
            sd.update(('film_generator.' + k, v.clone()) for k, v in self.film_gen_params.items())
        if self.classifier == 'versa':
            sd.update(('classifier.' + k, v.clone()) for k, v in self.versa_params.items())
        return sd

    def load_state_dict(self, sd):
        """Inverse of ``state_dict``: takes a reference-keyed checkpoint (e.g. the CUDA model's ``state_dict()``) so that
        both sides of a parity check hold the same weights. The FiLM gamma0/beta0 snapshot is re-taken from the loaded
        extractor, as the reference takes it at construction time (few_shot_recognisers.py:286)."""
        def sub(prefix):
            return OrderedDict((k[len(prefix):], v.detach().cpu().clone()) for k, v in sd.items() if k.startswith(prefix))
        self.extractor.load_state_dict(sub('feature_extractor.'), strict=True)
        if self.adapt_features:
            self.set_encoder_params = sub('set_encoder.')
            self.film_gen_params = sub('film_generator.')
            ext = dict(self.extractor.named_parameters())
            self.film_initial = {n: ext[n].detach().clone() for n in self.film_initial}
        if self.classifier == 'versa':
            self.versa_params = sub('classifier.')
        self.reset()

    # -- features -------------------------------------------------------------------------
    @torch.no_grad()
    def _features(self, clips, film=None):
        """Chunk by ``batch_size`` CLIPS, flatten to frames, backbone forward with the FiLM
        tensors substituted for the tagged norm affine params (functional_call), concatenate."""
        out = []
        for s in range(0, len(clips), self.batch_size):
            x = clips[s:s + self.batch_size]
            if x.dim() == 5:
                x = x.flatten(end_dim=1)
            x = x.float()
            out.append(functional_call(self.extractor, film, (x,)) if film else self.extractor(x))
        return torch.cat(out, dim=0)

    @torch.no_grad()
    def _task_embedding(self, clips):
        reps = [parts.set_encoder_forward(clips[s:s + self.batch_size].float(), self.set_encoder_params)
                for s in range(0, len(clips), self.batch_size)]
        return parts.task_embedding(reps)

    # -- single-step ----------------------------------------------------------------------
    @torch.no_grad()
    def personalise(self, context_clips, context_labels):
        context_labels = context_labels.cpu()
        if self.adapt_features:
            z = self._task_embedding(context_clips)
            self.film_dict, self.film_l2 = parts.film_generate(
                z, self.film_names, self.film_gen_params, self.film_initial)
        else:
            self.film_dict = {}
        f = parts.pool_clips(self._features(context_clips, self.film_dict), self.clip_length)
        self.context_features = f
        if self.classifier == 'proto':
            self.head = parts.proto_configure(f, context_labels, 'euclidean')
        elif self.classifier == 'proto_cosine':
            self.head = parts.proto_configure(f, context_labels, 'cosine')
        elif self.classifier == 'versa':
            self.head = parts.versa_configure(f, context_labels, self.versa_params)
        elif self.classifier == 'mahalanobis':
            self.head = parts.mahalanobis_configure(f, context_labels)
        else:
            raise ValueError("linear head is personalised with personalise_finetune()")

    @torch.no_grad()
    def predict(self, target_clips):
        if self.head is None:
            raise AttributeError("Weight and/or bias not set - is model personalised?")
        q = parts.pool_clips(self._features(target_clips, self.film_dict), self.clip_length)
        self.target_features = q
        s = self.logit_scale
        if self.classifier == 'proto':
            return parts.proto_predict(q, *self.head, s, 'euclidean')
        if self.classifier == 'proto_cosine':
            return parts.proto_predict(q, *self.head, s, 'cosine')
        if self.classifier in ('versa', 'linear'):
            return parts.linear_predict(q, *self.head, s)
        return parts.mahalanobis_predict(q, *self.head, s)

    # -- multi-step (FineTuner, frozen extractor) --------------------------------------------
    def personalise_finetune(self, context_clips, context_labels, num_grad_steps=50, learning_rate=1e-3,
                             optimizer='adam', betas=(0.9, 0.999), epsilon=1e-8, weight_decay=0.0,
                             momentum=0.0, recompute_features=False):
        """few_shot_recognisers.py:207-246 with the default FineTuner setting (frozen extractor,
        no FiLM): zero-init linear head (classifier_heads.py:59-60); per grad step, for each support
        batch: logits -> CE(mean) * batch_len/N -> backward; then ONE optimiser step.
        ``recompute_features=True`` re-runs the backbone every step exactly as the reference does
        (same numbers -- the extractor is frozen and in eval mode -- only slower)."""
        labels = context_labels.cpu().long()
        n = len(labels)
        num_classes = len(torch.unique(labels))
        w = torch.zeros(num_classes, self.feat_dim, requires_grad=True)
        b = torch.zeros(num_classes, requires_grad=True)
        if optimizer == 'adam':
            opt = torch.optim.Adam([w, b], lr=learning_rate, betas=betas, eps=epsilon, weight_decay=weight_decay)
        else:
            opt = torch.optim.SGD([w, b], lr=learning_rate, momentum=momentum, weight_decay=weight_decay)
        opt.zero_grad()
        feats = None if recompute_features else parts.pool_clips(self._features(context_clips), self.clip_length)
        for _ in range(num_grad_steps):
            for s in range(0, n, self.batch_size):
                if recompute_features:
                    x = parts.pool_clips(self._features(context_clips[s:s + self.batch_size]), self.clip_length)
                else:
                    x = feats[s:s + self.batch_size]
                y = labels[s:s + self.batch_size]
                loss = F.cross_entropy(parts.linear_predict(x, w, b, self.logit_scale), y)
                (loss * (len(y) / n)).backward()
            opt.step()
            opt.zero_grad()
        self.film_dict = {}
        self.head = (w.detach(), b.detach())

    def personalise_finetune_film(self, context_clips, context_labels, num_grad_steps=5, learning_rate=1e-3,
                                  optimizer='adam', betas=(0.9, 0.999), epsilon=1e-8, weight_decay=0.0, momentum=0.0):
        """FineTuner + FiLM: few_shot_recognisers.py:196-198 (the FiLM-tagged norm layers' weight / bias are unfrozen) and
        :207-246 (per grad step, per support batch: extractor forward in eval mode WITH autograd, pool, linear head,
        CE(mean) * batch_len / N, backward; then ONE optimiser step over the head and -- second parameter group, same
        learning rate: torch.optim ignores the 'lr_scale' key of utils/optim.py:27-30 -- the extractor's trainable
        parameters). Returns nothing; the extractor's FiLM parameters are updated in place, the head is stored."""
        labels = context_labels.cpu().long()
        n = len(labels)
        num_classes = len(torch.unique(labels))
        names = set(parts.film_parameter_names(self.name, self.extractor))
        film = []
        for pname, p in self.extractor.named_parameters():
            p.requires_grad_(pname in names)
            if pname in names:
                film.append(p)
        w = torch.zeros(num_classes, self.feat_dim, requires_grad=True)
        b = torch.zeros(num_classes, requires_grad=True)
        groups = [{'params': [w, b]}, {'params': film}]
        if optimizer == 'adam':
            opt = torch.optim.Adam(groups, lr=learning_rate, betas=betas, eps=epsilon, weight_decay=weight_decay)
        else:
            opt = torch.optim.SGD(groups, lr=learning_rate, momentum=momentum, weight_decay=weight_decay)
        opt.zero_grad()
        self.extractor.eval()
        for _ in range(num_grad_steps):
            for s in range(0, n, self.batch_size):
                clips = context_clips[s:s + self.batch_size]
                frames = clips.flatten(end_dim=1) if clips.dim() == 5 else clips
                x = parts.pool_clips(self.extractor(frames), self.clip_length)
                y = labels[s:s + self.batch_size]
                loss = F.cross_entropy(parts.linear_predict(x, w, b, self.logit_scale), y)
                (loss * (len(y) / n)).backward()
            opt.step()
            opt.zero_grad()
        for p in self.extractor.parameters():
            p.requires_grad_(False)
        self.film_dict = {}
        self.head = (w.detach(), b.detach())
