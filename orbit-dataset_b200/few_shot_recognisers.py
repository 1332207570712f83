"""Few-shot recognisers with the reference's Python API, running on liborbit_b200 (sm_100a).

Drop-in mirror of reference ``model/few_shot_recognisers.py`` (same class names, constructor
signatures, methods and attributes):
  FewShotRecogniser            :46-183
  MultiStepFewShotRecogniser   :185-269
  SingleStepFewShotRecogniser  :271-473
The control flow (batching by ``batch_size`` clips, FiLM generation, pooling, head configure /
predict, per-task state and reset) follows the reference; every tensor operation on the path is a
hand-written CUDA kernel reached through the C ABI.  There is no CPU fallback: tensors that are not
on an sm_100 device raise ``OrbitError``.
"""
import numpy as np
import torch
import torch.nn as nn

from . import lib as L
from .classifier_heads import ClassIndex, LinearClassifier, MeanPooler, PrototypicalClassifier
from .data_utils import get_batch_indices
from .feature_extractors import (create_feature_extractor, get_film_parameter_sizes, get_film_parameters,
                                 unfreeze_film)


class _HostStager:
    """Moves CPU-resident clips to the device on a side stream while the backbone runs (reference:
    ``clips.to(self.device, non_blocking=True)`` from pageable memory, few_shot_recognisers.py:112,142).

    One call = one whole-call device buffer (two are kept and alternate between calls), so the copy engine never
    waits for the backbone inside a call and the NEXT call's copy (e.g. the query set while the support set is still
    being processed) starts as soon as it is issued. The backbone consumes the buffer in chunks that only wait for
    their own bytes: a short ramp (96, 192, 320, 416 frames: the first kernels start after ~1.3 ms of copy and, at the
    measured 91 frames/ms of copy against ~75 frames/ms of backbone, no later pass waits for its bytes), then
    chunks of ``chunk_frames``. Pageable sources are staged through two pinned slices."""

    def __init__(self, device, chunk_frames, copy_frames=160):
        self.device = device
        self.chunk_frames = chunk_frames
        self.copy_frames = copy_frames          # granularity of the async copies (97 MB at 224 px)
        self.copy_stream = torch.cuda.Stream(device)
        self.dev = [None, None]                 # whole-call device buffers
        self.done = [None, None]                # the backbone finished reading dev[i]
        self.pinned = [None, None]
        self.pinned_free = [None, None]
        self.calls = 0
        self.bytes_copied = 0
        self.ramp = (96, 192, 320, 416)
        self.last_done = None                   # completion event of the previous call's last backbone pass
        self.last_buffer = None                 # (index, device view) of the most recent call's whole-call buffer

    def _plan(self, total, behind):
        # The ramp exists to start the backbone early on a call whose data is still on the host. A call issued while the
        # device is still working through the previous call's passes (predict() right after personalise(): the host
        # enqueues a pass in ~0.2 ms, the device needs ~18 us per frame, the copy engine ~11 us) finds its data
        # already on the device when its first kernel starts: no ramp, full-size passes. `behind` comes from the
        # previous call's completion EVENT (queued-work state), not from a timer, so the plan is reproducible.
        ramp = () if behind else self.ramp
        sizes, pos, k = [], 0, 0
        while pos < total:
            n = min(ramp[k], self.chunk_frames) if k < len(ramp) else self.chunk_frames
            n = min(n, total - pos)
            sizes.append(n)
            pos += n
            k += 1
        if len(sizes) > 1 and sizes[-1] < 64:   # no tiny tail pass: fold it into the previous one, or rebalance the
            tail = sizes.pop()                  # two when the engine would split the merged pass again
            if sizes[-1] + tail <= self.chunk_frames:
                sizes[-1] += tail
            else:
                both = sizes.pop() + tail
                first = min((both // 2 + 7) // 8 * 8, both - 1)
                sizes += [first, both - first]
        return sizes

    def stream(self, frames_cpu):
        """Yields device views of consecutive chunks of ``frames_cpu`` [F,3,H,W]; a view stays valid until the
        call after next."""
        compute = torch.cuda.current_stream(self.device)
        total = frames_cpu.shape[0]
        if total == 0:
            return
        i = self.calls % 2
        self.calls += 1
        per_frame = int(np.prod(frames_cpu.shape[1:]))
        need = total * per_frame
        if self.dev[i] is None or self.dev[i].numel() < need:
            if self.done[i] is not None:
                self.done[i].synchronize()
            self.dev[i] = torch.empty(max(need, self.chunk_frames * per_frame), dtype=torch.float32, device=self.device)
            # the allocator may hand back memory an earlier kernel on the compute stream still uses
            ev = torch.cuda.Event(); ev.record(compute); self.copy_stream.wait_event(ev)
        elif self.done[i] is not None:
            self.copy_stream.wait_event(self.done[i])
        dst = self.dev[i][:need].view(frames_cpu.shape)
        self.last_buffer = (i, dst)
        behind = self.last_done is not None and not self.last_done.query()
        sizes = self._plan(total, behind)
        pinned_src = frames_cpu.is_pinned()
        # Pinned source: every copy of the call is enqueued up front (the copy engine runs ahead of the backbone), then the
        # passes are handed out. Pageable source (what the reference's loaders deliver, data/queues.py:52): the host-side
        # staging memcpy of a pass runs while the device works on the previous pass, so a pass is handed out as soon as ITS
        # copies are enqueued (staging everything first left the device idle for the whole 25 ms of host copying).
        events, pos, j = [], 0, 0
        for n in sizes:
            start, end = pos, pos + n
            while pos < end:
                m = min(self.copy_frames, end - pos)
                src = frames_cpu[pos:pos + m]
                if not pinned_src:
                    k = j % 2
                    j += 1
                    if self.pinned[k] is None or self.pinned[k].numel() < m * per_frame:
                        self.pinned[k] = torch.empty(self.copy_frames * per_frame, dtype=torch.float32).pin_memory()
                        self.pinned_free[k] = None
                    if self.pinned_free[k] is not None:
                        self.pinned_free[k].synchronize()   # staging slice still in flight
                    stage = self.pinned[k][:m * per_frame].view(src.shape)
                    stage.copy_(src)
                    src = stage
                with torch.cuda.stream(self.copy_stream):
                    dst[pos:pos + m].copy_(src, non_blocking=True)
                    if not pinned_src:
                        self.pinned_free[k] = torch.cuda.Event()
                        self.pinned_free[k].record(self.copy_stream)
                pos += m
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
            if pinned_src:
                events.append(ev)
            else:
                self.bytes_copied += n * per_frame * 4
                compute.wait_event(ev)
                yield dst[start:end]
        if pinned_src:
            self.bytes_copied += need * 4
            pos = 0
            for n, ev in zip(sizes, events):
                compute.wait_event(ev)
                yield dst[pos:pos + n]
                pos += n
        self.release(i)

    def release(self, i):
        """Marks buffer ``i`` as consumed up to this point of the compute stream (call again after enqueuing more work
        that reads the buffer, e.g. a second traversal of the same staged clips)."""
        self.done[i] = torch.cuda.Event()
        self.done[i].record(torch.cuda.current_stream(self.device))
        self.last_done = self.done[i]


class FewShotRecogniser(nn.Module):
    """Generic few-shot classification model (few_shot_recognisers.py:46-183)."""

    def __init__(self, feature_extractor_name: str, adapt_features: bool, classifier: str, clip_length: int,
                 batch_size: int, learn_extractor: bool, logit_scale: float = 1.0):
        super().__init__()
        self.adapt_features = adapt_features
        self.learn_extractor = learn_extractor
        self.clip_length = clip_length
        self.batch_size = batch_size
        self.logit_scale = logit_scale

        self.feature_extractor, self.film_parameter_names = create_feature_extractor(
            feature_extractor_name=feature_extractor_name, pretrained=True, with_film=self.adapt_features,
            learn_extractor=self.learn_extractor)

        self.classifier_name = classifier
        if classifier == 'linear':
            self.classifier = LinearClassifier(self.feature_extractor.output_size, self.logit_scale)
        elif classifier == 'versa':
            from .classifier_heads_ext import VersaClassifier
            self.classifier = VersaClassifier(self.feature_extractor.output_size, self.logit_scale)
        elif classifier == 'proto':
            self.classifier = PrototypicalClassifier(self.logit_scale)
        elif classifier == 'proto_cosine':
            self.classifier = PrototypicalClassifier(self.logit_scale, distance_fn='cosine')
        elif classifier == 'mahalanobis':
            from .classifier_heads_ext import MahalanobisClassifier
            self.classifier = MahalanobisClassifier(self.logit_scale)
        else:
            raise ValueError(f"Classifier {classifier} not valid.")

        self.frame_pooler = MeanPooler(T=self.clip_length)
        # 'reference': OpsCounter totals equal the reference's (which re-runs the frozen extractor in every FineTuner
        # grad step, few_shot_recognisers.py:231-246); 'actual': the MACs this library really executes
        self.mac_accounting = 'reference'
        self.device = torch.device('cpu')
        self._stager = None
        self.stage_copy_frames = 160      # frames per async H2D copy for CPU-resident clips (97 MB at 224 px)
        # sizes of the first backbone passes of a call (frames). Tuned on B200 with the measured pass times (scripts/e2e_probe.py,
        # scripts/pass_overhead.py): the copy of a 1,600-frame support set ends at 17.4 ms and the LAST pass can only start then, so
        # it should be short (640), while earlier passes must not be so small that their fixed cost dominates: 39.1 -> 37.7 ms
        self.stage_ramp = (96, 192, 320, 416)

    def _set_device(self, device):
        self.device = torch.device(device)

    def _send_to_device(self):
        self.to(self.device)

    # ---- features --------------------------------------------------------------------------------
    def _require_device(self):
        if self.device.type != 'cuda':
            raise L.OrbitError("orbit_b200 runs only on a CUDA (sm_100) device; call _set_device('cuda:N') and "
                               "_send_to_device() first (there is no CPU fallback)")

    def _film_blob(self, film_dict):
        return None  # overridden by SingleStepFewShotRecogniser

    def _host_stager(self):
        chunk = self.feature_extractor.get_option('chunk_frames')
        if self._stager is None or self._stager.device != self.device or self._stager.chunk_frames != chunk:
            self._stager = _HostStager(self.device, chunk)
        self._stager.copy_frames, self._stager.ramp = self.stage_copy_frames, tuple(self.stage_ramp)
        return self._stager

    def _run_extractor(self, frames, film_dict):
        """frames [F,3,H,W] on CPU or device -> [F, D] on device."""
        blob = self._film_blob(film_dict) if film_dict else None
        if not frames.is_cuda and blob is not None and blob.requires_grad and torch.is_grad_enabled():
            # meta-training: the frames are kept for the backward pass, so they get their own device copy (the stager's
            # buffers are recycled by the next call)
            frames = frames.to(self.device, non_blocking=True).float()
        if frames.is_cuda:
            return self.feature_extractor(frames, blob)
        # Host clips arrive in passes (a short ramp, then chunk_frames), each waiting only for its own bytes. (Tried in round 2
        # and removed: alternating consecutive passes between two streams / workspaces so that the tail of one pass overlaps
        # the head of the next -- no gain, the persistent GEMM kernels own every SM's shared memory)
        outs = [self.feature_extractor(dev_frames, blob) for dev_frames in self._host_stager().stream(frames.float())]
        if not outs:
            return torch.empty(0, self.feature_extractor.output_size, device=self.device)
        return outs[0] if len(outs) == 1 else torch.cat(outs, dim=0)

    def _get_features(self, clips, film_dict={}, ops_counter=None):
        """few_shot_recognisers.py:99-122."""
        self._require_device()
        if len(clips.shape) == 5:
            clips = clips.reshape(clips.shape[0] * clips.shape[1], *clips.shape[2:])
        if ops_counter:
            ops_counter.compute_macs(self.feature_extractor, clips)
        return self._run_extractor(clips, film_dict)

    def _get_features_in_batches(self, clips, film_dict={}, ops_counter=None):
        """few_shot_recognisers.py:124-153: chunks of ``batch_size`` CLIPS, flattened to frames."""
        self._require_device()
        features = []
        num_clips = len(clips)
        num_batches = int(np.ceil(float(num_clips) / float(self.batch_size)))
        for batch in range(num_batches):
            batch_start_index, batch_end_index = get_batch_indices(batch, num_clips, self.batch_size)
            batch_clips = clips[batch_start_index:batch_end_index]
            if len(batch_clips.shape) == 5:
                batch_clips = batch_clips.flatten(end_dim=1)
            features.append(self._run_extractor(batch_clips, film_dict))
            if ops_counter:
                ops_counter.compute_macs(self.feature_extractor, batch_clips)
        if not features:
            return torch.empty(0, self.feature_extractor.output_size, device=self.device)
        return features[0] if len(features) == 1 else torch.cat(features, dim=0)

    def _pool_features(self, features, ops_counter=None):
        """few_shot_recognisers.py:155-166."""
        if ops_counter:
            ops_counter.add_macs(features.size(0) * features.size(1))
        return self.frame_pooler(features)

    def predict_video(self, video_frames, want_argmax=False):
        """Per-frame logits of one target video: equal (bit for bit) to
        ``predict(attach_frame_history(video_frames, clip_length))`` -- the reference's test loop,
        single-step-learner.py:327-332 -- but every frame goes through the extractor ONCE instead of ``clip_length``
        times; the causal ``clip_length``-frame means are formed on the device from the per-frame features
        (valid because the extractor is frame-wise and in eval mode). ``video_frames``: [F,3,H,W] on CPU or device."""
        self._set_batch_norm_state()
        self._require_device()
        if video_frames.dim() != 4:
            raise ValueError("predict_video expects the frames of one video, [F,3,H,W]")
        film_dict = getattr(self, 'film_dict', None) or {}
        step = self.batch_size * self.clip_length
        feats = [self._run_extractor(video_frames[i:i + step], film_dict) for i in range(0, len(video_frames), step)]
        if not feats:
            feats = [torch.empty(0, self.feature_extractor.output_size, device=self.device)]
        feats = feats[0] if len(feats) == 1 else torch.cat(feats, dim=0)
        if self.clip_length > 1 and len(feats):
            pooled = torch.empty_like(feats)
            L.check(L.load().orbit_pool_history(L.ptr(feats), feats.shape[0], self.clip_length, feats.shape[1], L.ptr(pooled),
                                                L.stream_ptr(feats.device)), "orbit_pool_history")
            L.count_launches(1)
            feats = pooled
        kwargs = {'want_argmax': True} if want_argmax else {}
        return self.classifier.predict(feats, clip_length=1, **kwargs)

    def set_test_mode(self, test_mode):
        self.test_mode = test_mode

    def _set_batch_norm_state(self):
        """few_shot_recognisers.py:176-183. The native extractor always normalises with the running
        statistics (eval mode); meta-training an unfrozen extractor with batch statistics is not
        implemented and is refused rather than silently approximated."""
        self.eval()
        if self.learn_extractor and not self.test_mode:
            raise NotImplementedError("training the extractor's own weights (train-mode BatchNorm + weight gradients) is "
                                      "outside the B200 hot path implemented so far (SURVEY.md 8f-3)")
        if isinstance(self, SingleStepFewShotRecogniser) and self.adapt_features:
            # CNAPs-style meta-training (single-step-learner.py:196-243): outside test mode and with autograd on, the set
            # encoder / FiLM generator / extractor / head calls build a graph of native backward kernels (training.py)
            self.set_encoder.train_graph = self._meta_training()

    def _meta_training(self):
        return bool(self.adapt_features and not getattr(self, 'test_mode', True) and torch.is_grad_enabled())


class SingleStepFewShotRecogniser(FewShotRecogniser):
    """ProtoNets / CNAPs / SimpleCNAPs: personalised in one forward pass (few_shot_recognisers.py:271-473)."""

    def __init__(self, feature_extractor_name: str, adapt_features: bool, classifier: str, clip_length: int,
                 batch_size: int, learn_extractor: bool, num_lite_samples: int, logit_scale: float = 1.0):
        super().__init__(feature_extractor_name, adapt_features, classifier, clip_length, batch_size, learn_extractor,
                         logit_scale)
        self.num_lite_samples = num_lite_samples
        if self.adapt_features:
            from .feature_adapters import FilmParameterGenerator, SetEncoder
            self.set_encoder = SetEncoder()
            self.film_parameter_sizes = get_film_parameter_sizes(self.film_parameter_names, self.feature_extractor)
            initial_film_parameters = get_film_parameters(self.film_parameter_names, self.feature_extractor)
            self.film_generator = FilmParameterGenerator(self.film_parameter_sizes, initial_film_parameters,
                                                         pooled_size=self.set_encoder.output_size,
                                                         hidden_size=self.set_encoder.output_size)
        else:
            from .feature_adapters import NullGenerator, NullSetEncoder
            self.set_encoder = NullSetEncoder()
            self.film_generator = NullGenerator()
        self.film_dict = None
        self._staged_context = None
        self.test_mode = False

    def _reset(self):
        self.film_dict = None
        self.classifier.reset()

    def _clear_caches(self):
        self.reps_cache = None
        self.features_cache = None

    def _film_blob(self, film_dict):
        return self.film_generator.as_blob(film_dict)

    def personalise(self, context_clips, context_labels, ops_counter=None):
        """few_shot_recognisers.py:313-326."""
        self._set_batch_norm_state()
        self._require_device()
        # the task's only host sync (label values -> class count), taken before anything is enqueued: see ClassIndex
        class_index = ClassIndex(context_labels, self.device)
        self._staged_context = None
        task_embedding = self._get_task_embedding_in_batches(context_clips, ops_counter)
        self.film_dict = self._generate_film_params(task_embedding, ops_counter)
        if self._meta_training():
            # the support -> head path carries no gradient (configure() makes leaf Parameters): inference kernels
            with torch.no_grad():
                context_features = self._get_features_in_batches(context_clips, self.film_dict, ops_counter)
        elif self._staged_context is not None:
            # CPU-resident clips the set encoder has already pulled onto the device: the extractor reads that copy
            # (the reference copies the support set H2D twice, few_shot_recognisers.py:322-324 / SURVEY Appendix B-6)
            buffer_index, context_dev = self._staged_context
            context_features = self._get_features_in_batches(context_dev, self.film_dict, ops_counter)
            self._stager.release(buffer_index)
            self._staged_context = None
        else:
            context_features = self._get_features_in_batches(context_clips, self.film_dict, ops_counter)
        # pooling (poolers.py:13-16) is fused into the head's configure kernel; its MACs are still the reference's
        if ops_counter:
            ops_counter.add_macs(context_features.size(0) * context_features.size(1))
        self.classifier.configure(context_features, context_labels, ops_counter, clip_length=self.clip_length,
                                  class_index=class_index)

    def personalise_with_lite(self, context_clips, context_labels):
        """LITE (few_shot_recognisers.py:328-343): a random subset of ``num_lite_samples`` context clips is processed with
        back-propagation enabled, the remainder comes from a no-grad cache. ``np.random.permutation`` is unseeded here as in
        the reference."""
        self._set_batch_norm_state()
        self._require_device()
        shuffled_idxs = np.random.permutation(len(context_clips))
        grad_idxs = shuffled_idxs[0:self.num_lite_samples]
        no_grad_idxs = shuffled_idxs[self.num_lite_samples:]
        self._staged_context = None
        task_embedding = self._get_task_embedding_with_split_batch(context_clips, grad_idxs, no_grad_idxs)
        self.film_dict = self._generate_film_params(task_embedding)
        context_features = self._get_features_with_split_batch(context_clips, self.film_dict, grad_idxs, no_grad_idxs)
        shuffled_labels = context_labels[torch.as_tensor(shuffled_idxs, device=context_labels.device)]
        self.classifier.configure(context_features, shuffled_labels, clip_length=self.clip_length)

    def _get_task_embedding(self, context_clips, ops_counter=None, aggregation='mean'):
        """few_shot_recognisers.py:345-359."""
        context_clips = context_clips.to(self.device, non_blocking=True)
        reps = self.set_encoder(context_clips)
        if ops_counter:
            ops_counter.compute_macs(self.set_encoder, context_clips)
        return self.set_encoder.aggregate(reps, aggregation=aggregation)

    def _get_task_embedding_with_split_batch(self, context_clips, grad_idxs, no_grad_idxs):
        """few_shot_recognisers.py:388-413. The cache holds PER-FRAME embeddings and is indexed with CLIP indices, exactly as
        the reference does (identical for clip_length 1, its LITE training setting)."""
        from .feature_adapters import NullSetEncoder
        if isinstance(self.set_encoder, NullSetEncoder):
            return None
        self._set_batch_norm_state()
        if getattr(self, 'reps_cache', None) is None:
            with torch.set_grad_enabled(False):
                self.reps_cache = self._get_task_embedding_in_batches(context_clips, aggregation='none')
        with torch.set_grad_enabled(True):
            reps_with_grads = self._get_task_embedding(context_clips[grad_idxs], aggregation='none')
        reps_without_grads = self.reps_cache[torch.as_tensor(no_grad_idxs, device=self.reps_cache.device)]
        return torch.cat((reps_with_grads, reps_without_grads)).mean(dim=0)

    def _get_features_with_split_batch(self, context_clips, film_dict, grad_idxs, no_grad_idxs):
        """few_shot_recognisers.py:415-437. With a frozen extractor no gradient flows from the head back into the context
        features (the head's weights are leaf Parameters): the ``with grads`` part is computed by the same
        inference kernels (equal values), and the returned rows are ordered [grad_idxs, no_grad_idxs] as in the reference."""
        self._set_batch_norm_state()
        with torch.no_grad():
            if getattr(self, 'features_cache', None) is None:
                if self._staged_context is not None:      # the set-encoder cache pass already staged the host clips
                    buffer_index, context_dev = self._staged_context
                    self.features_cache = self._get_features_in_batches(context_dev, film_dict)
                    self._stager.release(buffer_index)
                else:
                    self.features_cache = self._get_features_in_batches(context_clips, film_dict)
            self._staged_context = None
            features_with_grads = self._get_features(context_clips[grad_idxs], film_dict)
            features_without_grads = self.features_cache[torch.as_tensor(no_grad_idxs, device=self.features_cache.device)]
            return torch.cat((features_with_grads, features_without_grads))

    def _get_task_embedding_in_batches(self, context_clips, ops_counter=None, aggregation='mean'):
        """few_shot_recognisers.py:361-386."""
        from .feature_adapters import NullSetEncoder
        if isinstance(self.set_encoder, NullSetEncoder):
            return None
        self._require_device()
        reps = []
        num_clips = len(context_clips)
        if not context_clips.is_cuda and num_clips and not (self.set_encoder.train_graph and torch.is_grad_enabled()):
            # host clips: ONE staged H2D of the whole support set (pinned or pageable source, side stream); the set
            # encoder consumes it pass by pass and personalise() hands the same device copy to the extractor. The
            # per-frame embeddings do not depend on how frames are grouped into passes, so batch_size only matters
            # for the MAC accounting below.
            stager = self._host_stager()
            frames = context_clips.flatten(end_dim=1) if context_clips.dim() == 5 else context_clips
            reps = [self.set_encoder(dev_frames) for dev_frames in stager.stream(frames.float())]
            buffer_index, staged = stager.last_buffer
            self._staged_context = (buffer_index, staged.view(context_clips.shape))
            if ops_counter:
                ops_counter.compute_macs(self.set_encoder, context_clips)
            return self.set_encoder.aggregate(reps, aggregation=aggregation)
        num_batches = int(np.ceil(float(num_clips) / float(self.batch_size)))
        for batch in range(num_batches):
            batch_start_index, batch_end_index = get_batch_indices(batch, num_clips, self.batch_size)
            batch_clips = context_clips[batch_start_index:batch_end_index].to(self.device, non_blocking=True)
            reps.append(self.set_encoder(batch_clips))
            if ops_counter:
                ops_counter.compute_macs(self.set_encoder, batch_clips)
        return self.set_encoder.aggregate(reps, aggregation=aggregation)

    def _generate_film_params(self, task_embedding, ops_counter=None):
        """few_shot_recognisers.py:439-451."""
        if ops_counter:
            ops_counter.compute_macs(self.film_generator, task_embedding)
        return self.film_generator(task_embedding)

    def predict(self, target_clips, want_argmax=False):
        """few_shot_recognisers.py:453-462 (+ optional fused arg-max output)."""
        self._set_batch_norm_state()
        target_features = self._get_features_in_batches(target_clips, self.film_dict)
        return self.classifier.predict(target_features, clip_length=self.clip_length, **(
            {'want_argmax': True} if want_argmax else {}))

    def predict_a_batch(self, target_clips):
        """few_shot_recognisers.py:464-473."""
        self._set_batch_norm_state()
        target_features = self._get_features(target_clips, self.film_dict)
        return self.classifier.predict(target_features, clip_length=self.clip_length)


class MultiStepFewShotRecogniser(FewShotRecogniser):
    """FineTuner: personalised with gradient steps on a new linear head (few_shot_recognisers.py:185-269)."""

    def __init__(self, feature_extractor_name: str, adapt_features: bool, classifier: str, clip_length: int,
                 batch_size: int, learn_extractor: bool, logit_scale: float = 1.0):
        super().__init__(feature_extractor_name, adapt_features, classifier, clip_length, batch_size, learn_extractor,
                         logit_scale)
        if self.adapt_features:
            self.film_parameter_sizes = get_film_parameter_sizes(self.film_parameter_names, self.feature_extractor)
            unfreeze_film(self.film_parameter_names, self.feature_extractor)
        self.test_mode = True

    def _reset(self):
        self.classifier.reset()

    def init_classifier(self, num_classes: int):
        self.classifier.init(num_classes)
        self.classifier.to(self.device)

    def personalise(self, context_clips, context_labels, learning_args, ops_counter=None):
        """few_shot_recognisers.py:207-246 for the default FineTuner (frozen extractor, new linear head).
        The extractor is frozen and in eval mode, so the support features are loop-invariant: they are
        computed ONCE, then the num_grad_steps x batches loop (CE mean * batch_len/N, one optimiser step
        per grad step) runs on the device in a single kernel."""
        from .finetune import finetune_linear_head
        self._set_batch_norm_state()
        num_grad_steps = learning_args.pop('num_grad_steps')
        learning_rate = learning_args.pop('learning_rate')
        optimizer = learning_args.pop('optimizer')
        learning_args.pop('loss_fn')
        learning_args.pop('extractor_lr_scale')
        if self.learn_extractor:
            raise NotImplementedError("fine-tuning the extractor's weights needs weight-gradient kernels (SURVEY.md 8f-3)")
        num_classes = len(torch.unique(context_labels))
        self.init_classifier(num_classes)
        if self.adapt_features:
            return self._personalise_film(context_clips, context_labels, num_grad_steps, learning_rate, optimizer,
                                          dict(learning_args), ops_counter)
        macs_before = ops_counter.get_task_macs() if ops_counter else 0
        features = self._get_features_in_batches(context_clips, ops_counter=ops_counter)
        features = self._pool_features(features, ops_counter=ops_counter)
        if ops_counter:
            # The frozen features are computed ONCE here; the reference recomputes extractor + pooling in every grad
            # step (few_shot_recognisers.py:231-246). "MACs to personalise" is a reported ORBIT metric, so by default
            # the counter gets the reference's figure (x num_grad_steps); mac_accounting = 'actual' keeps what ran.
            if self.mac_accounting == 'reference' and num_grad_steps > 1:
                ops_counter.add_macs((num_grad_steps - 1) * (ops_counter.get_task_macs() - macs_before))
            # the head's forward inside the loop (classifier_heads.py:72-73), once per clip per grad step
            ops_counter.add_macs(num_grad_steps * num_classes * features.size(0) * features.size(1))
        finetune_linear_head(self.classifier, features, context_labels, self.batch_size, num_grad_steps,
                             learning_rate, optimizer, dict(learning_args), self.logit_scale)

    def _personalise_film(self, context_clips, context_labels, num_grad_steps, learning_rate, optimizer, opt_args, ops_counter):
        """FineTuner + FiLM (few_shot_recognisers.py:196-198,225-246): gradient steps on the linear head AND on the affine
        weight / bias of the FiLM-tagged BatchNorms, through the frozen extractor (BatchNorm in eval mode). Per batch: a
        forward that keeps pre-activations, the head + cross-entropy backward, the extractor backward (native kernels,
        csrc/train.cu + the tcgen05 GEMM on transposed weights); per grad step: one optimiser step (utils/optim.py:11-32;
        torch.optim over the ~20 k trainable values, exactly the reference's optimiser)."""
        import ctypes as C
        from .classifier_heads import _class_index
        lib = L.load()
        self._require_device()
        fe, dev = self.feature_extractor, self.device
        classes, idx = _class_index(context_labels)
        labels_dev = torch.from_numpy(idx).to(dev)
        film_params = [p for n, p in fe.named_parameters() if n in set(self.film_parameter_names)]
        head_params = [self.classifier.weight, self.classifier.bias]
        if optimizer == 'adam':
            opt = torch.optim.Adam([{'params': head_params}, {'params': film_params}], lr=learning_rate,
                                   eps=opt_args.get('epsilon', 1e-8), weight_decay=opt_args.get('weight_decay', 0.0),
                                   betas=tuple(opt_args.get('betas', (0.9, 0.999))))
        elif optimizer == 'sgd':
            opt = torch.optim.SGD([{'params': head_params}, {'params': film_params}], lr=learning_rate,
                                  momentum=opt_args.get('momentum', 0.0), weight_decay=opt_args.get('weight_decay', 0.0))
        else:
            raise ValueError(f"Optimizer {optimizer} not valid.")
        n_ctx, Lc, d, c = len(context_labels), self.clip_length, fe.output_size, len(classes)
        num_batches = int(np.ceil(float(n_ctx) / float(self.batch_size)))
        gw, gb = torch.zeros_like(self.classifier.weight), torch.zeros_like(self.classifier.bias)
        self.classifier.weight.grad, self.classifier.bias.grad = gw, gb
        scratch = torch.empty(lib.orbit_linear_ce_scratch_floats(self.batch_size, d, c), dtype=torch.float32, device=dev)
        for _ in range(num_grad_steps):
            for batch in range(num_batches):
                b0, b1 = get_batch_indices(batch, n_ctx, self.batch_size)
                clips = context_clips[b0:b1]
                frames = clips.flatten(end_dim=1) if clips.dim() == 5 else clips
                feats = fe.forward_train(frames.to(dev, non_blocking=True).float())
                if ops_counter:
                    ops_counter.compute_macs(fe, frames)
                    ops_counter.add_macs(feats.size(0) * feats.size(1))
                    ops_counter.add_macs(c * (b1 - b0) * d)
                dfeats = torch.empty_like(feats)
                L.check(lib.orbit_linear_ce_backward(L.ptr(feats), L.ptr(labels_dev[b0:b1].contiguous()), L.ptr(self.classifier.weight.detach()),
                                                     L.ptr(self.classifier.bias.detach()), b1 - b0, Lc, d, c, float(self.logit_scale),
                                                     float(b1 - b0) / float(n_ctx), L.ptr(gw), L.ptr(gb), L.ptr(dfeats), L.ptr(scratch),
                                                     L.stream_ptr(dev)), "orbit_linear_ce_backward")
                L.count_launches(2)
                fe.backward_train(dfeats)
            opt.step()
            gw.zero_(); gb.zero_()
            fe.zero_film_grads()
        for p in film_params + head_params:
            p.grad = None
        fe._grad_blob = None

    def predict(self, clips, ops_counter=None):
        """few_shot_recognisers.py:248-258."""
        self._set_batch_norm_state()
        features = self._get_features_in_batches(clips, ops_counter=ops_counter)
        if ops_counter:
            ops_counter.add_macs(features.size(0) * features.size(1))      # pooling, fused into the head kernel
        return self.classifier.predict(features, ops_counter=ops_counter, clip_length=self.clip_length)

    def personalise_with_lite(self, context_clips, context_labels):
        NotImplementedError
