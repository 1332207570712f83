"""Autograd bridges for meta-training with a frozen extractor.

The reference gets its gradients from torch autograd (single-step-learner.py:196-243: ``loss.backward()`` through
``predict`` -> extractor with ``film_dict`` -> ``FilmParameterGenerator`` -> ``SetEncoder``). Here every differentiable
stage of that chain is a ``torch.autograd.Function`` whose forward AND backward are native kernels of liborbit_b200
(csrc/train.cu, csrc/train_setenc.cu, the tcgen05 GEMM on transposed weights); autograd only routes the tensors:

    support clips --SetEncoderFn--> per-frame embeddings --mean--> z --FilmGeneratorFn--> film blob (gamma', beta')
    query clips, film blob --ExtractorFilmFn--> frame features --HeadPredictFn--> logits

The support -> head path carries no gradient, exactly as in the reference (configure() wraps the head weights in
``nn.Parameter``: classifier_heads.py:179-180,261-263), so the support pass runs the inference kernels.
"""
import torch

from . import lib as L


def _param_views(grad_blob, module, params):
    """grads of ``params`` (a tuple of nn.Parameters of ``module``) as views of a gradient blob laid out like the module's blob"""
    by_id = module._param_slots()
    out = []
    for p in params:
        off, numel, shape = by_id[id(p)]
        out.append(grad_blob[off:off + numel].view(shape))
    return tuple(out)


class ExtractorFilmFn(torch.autograd.Function):
    """frames [B,3,H,W], film blob -> features [B,D]; backward: d loss / d film blob (few_shot_recognisers.py:114-117)."""

    @staticmethod
    def forward(ctx, extractor, frames, film_blob, generation):
        film = film_blob.detach()
        film._orbit_generation = generation          # the prepare() cache key of feature_extractors.py
        feats, state = extractor._forward_train_impl(frames, film, fresh_arena=True)
        ctx.extractor, ctx.state, ctx.film = extractor, state, film
        return feats

    @staticmethod
    def backward(ctx, dfeats):
        fe = ctx.extractor
        lib = L.load()
        grad_blob = torch.zeros(fe._n_floats, dtype=torch.float32, device=dfeats.device)
        fe._backward_train_impl(dfeats, ctx.state, ctx.film, grad_blob)
        gfilm = torch.empty_like(ctx.film)
        L.check(lib.orbit_engine_film_grad(fe._engine, L.ptr(grad_blob), L.ptr(gfilm), L.stream_ptr(dfeats.device)),
                "orbit_engine_film_grad")
        ctx.state = None
        return None, None, gfilm, None


class SetEncoderFn(torch.autograd.Function):
    """frames -> per-frame 64-d embeddings; backward: gradients of every set-encoder parameter (set_encoders.py:81-120)."""

    @staticmethod
    def forward(ctx, encoder, frames, *params):
        feats, state = encoder._forward_train_impl(frames, None, fresh_arena=True)
        ctx.encoder, ctx.state, ctx.params = encoder, state, params
        return feats

    @staticmethod
    def backward(ctx, dfeats):
        enc = ctx.encoder
        grad_blob = torch.zeros(enc._n_floats, dtype=torch.float32, device=dfeats.device)
        enc._backward_train_impl(dfeats, ctx.state, None, grad_blob)
        ctx.state = None
        return (None, None) + _param_views(grad_blob, enc, ctx.params)


class FilmGeneratorFn(torch.autograd.Function):
    """task embedding z [hidden] -> film blob; backward: generator parameter gradients + dz (feature_adapters.py:66-78)."""

    @staticmethod
    def forward(ctx, generator, z, *params):
        film = generator._generate(z.detach())
        ctx.generator, ctx.z, ctx.params = generator, z.detach(), params
        return film

    @staticmethod
    def backward(ctx, dfilm):
        gen = ctx.generator
        lib = L.load()
        dev = dfilm.device
        grad_blob = torch.zeros_like(gen._blob)
        dz = torch.empty(gen.hidden, dtype=torch.float32, device=dev)
        scratch = torch.empty(len(gen._rows) * gen.hidden, dtype=torch.float32, device=dev)
        z = ctx.z.reshape(-1).contiguous().float()
        L.check(lib.orbit_film_generate_backward(L.ptr(gen._blob), L.ptr(gen._table_dev), len(gen._rows), L.ptr(z), gen.hidden,
                                                 L.ptr(dfilm.contiguous()), L.ptr(grad_blob), L.ptr(dz), L.ptr(scratch),
                                                 L.stream_ptr(dev)), "orbit_film_generate_backward")
        L.count_launches(2)
        return (None, dz.view(ctx.z.shape)) + _param_views(grad_blob, gen, ctx.params)


class HeadPredictFn(torch.autograd.Function):
    """frame features [N*L, D] -> logits [N, C] for the linear-form heads; backward: d loss / d frame features (the head's
    weights are leaf Parameters created by configure(): they receive no gradient in the reference either)."""

    @staticmethod
    def forward(ctx, features, clip_length, weight, bias, metric, logit_scale):
        from .classifier_heads import _head_predict_raw
        features = features.detach().contiguous().float()
        ctx.save_for_backward(features, weight)
        ctx.args = (clip_length, metric, logit_scale)
        return _head_predict_raw(features, clip_length, weight, bias, metric, logit_scale, False)

    @staticmethod
    def backward(ctx, dlogits):
        features, weight = ctx.saved_tensors
        clip_length, metric, logit_scale = ctx.args
        n = features.shape[0] // clip_length
        c, d = weight.shape
        dfeat = torch.empty_like(features)
        L.check(L.load().orbit_head_predict_backward(L.ptr(features), L.ptr(weight), L.ptr(dlogits.contiguous().float()), n, clip_length,
                                                     d, c, metric, float(logit_scale), L.ptr(dfeat), L.stream_ptr(features.device)),
                "orbit_head_predict_backward")
        L.count_launches(1)
        return dfeat, None, None, None, None, None


class MahalanobisPredictFn(torch.autograd.Function):
    """frame features -> Mahalanobis logits (classifier_heads.py:328-350); backward: d loss / d frame features. The class
    means / precisions are leaf Parameters created by configure() (classifier_heads.py:323-326): no gradient, as in the reference."""

    @staticmethod
    def forward(ctx, head, features, clip_length):
        features = features.detach().contiguous().float()
        ctx.save_for_backward(features, head.means.detach(), head.precisions.detach())
        ctx.args = (clip_length, head.logit_scale)
        return head._predict_raw(features, clip_length)

    @staticmethod
    def backward(ctx, dlogits):
        features, means, precisions = ctx.saved_tensors
        clip_length, logit_scale = ctx.args
        n = features.shape[0] // clip_length
        c, d = means.shape
        dfeat = torch.empty_like(features)
        L.check(L.load().orbit_mahalanobis_predict_backward(L.ptr(features), L.ptr(means), L.ptr(precisions.contiguous()),
                                                            L.ptr(dlogits.contiguous().float()), n, clip_length, d, c, float(logit_scale),
                                                            L.ptr(dfeat), L.stream_ptr(features.device)), "orbit_mahalanobis_predict_backward")
        L.count_launches(1)
        return None, dfeat, None


def allreduce_gradients(parameters, group=None):
    """Data-parallel meta-training (training variant): the ``tasks_per_batch`` tasks of one optimiser step
    (single-step-learner.py:150-166) are dealt round-robin to the ranks (``evaluation.shard_episodes``), every rank
    accumulates the gradients of its own tasks -- each task's loss is already divided by ``tasks_per_batch``
    (single-step-learner.py:203), so the per-rank gradients simply ADD -- and this sums them over the ranks with ONE
    all-reduce (NCCL over NVLink on the GPU box, gloo on CPU) of one flat buffer before ``optimizer.step()``.
    Parameters without a gradient on this rank (no task dealt to it) contribute zeros. Returns the number of values reduced."""
    import torch.distributed as dist
    params = [p for p in parameters if p.requires_grad]
    if not params:
        return 0
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).float() for p in params])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    offset = 0
    for p in params:
        n = p.numel()
        g = flat[offset:offset + n].view_as(p)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        offset += n
    return offset
