"""Throughput of the other BASELINE.json configurations on one B200 (device-resident synthetic episodes, CUDA events).
They are parity-test cases, not bench lines (bench.py measures configs[1]); this records what the same kernels deliver
on them.  S1: ProtoNet+resnet18 84px 5-way 5-shot;  S3: CNAPs (versa + FiLM) resnet18 224px, 5..15-way, 10-shot, 20 query
clips per class;  S4: FineTuner vit_b_32 224px 8-way 10-shot, 50 Adam steps, 160 query frames."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import orbit_b200
from orbit_b200.synthetic import EpisodeSpec, calibration_frames, make_episode

dev = torch.device('cuda:0')


def prepare(m, size):
    m._set_device(dev); m._send_to_device(); m.set_test_mode(True)
    fe = m.feature_extractor
    fe.reset_parameters(1991)
    try:
        fe.calibrate_batchnorm(calibration_frames(size).to(dev))
    except Exception:
        pass                                    # ViT: no BatchNorm
    if getattr(m, 'adapt_features', False) and hasattr(m, 'film_generator'):
        from orbit_b200.feature_extractors import get_film_parameters
        m.film_generator.initial_film_parameters = get_film_parameters(m.film_parameter_names, fe)
    return m


def timed(step, n_warm, n):
    for i in range(n_warm): step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): step(n_warm + i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def run(name, model, episodes, personalise, n=12):
    eps = [tuple(t.to(dev) for t in e) for e in episodes]
    def step(i):
        c, cy, t, ty = eps[i % len(eps)]
        personalise(model, c, cy)
        logits = model.predict(t)
        model._reset()
        return logits
    ms = timed(step, 2, n)
    qf = sum(e[2].shape[0] * e[2].shape[1] for e in episodes) / len(episodes)
    print(json.dumps({"config": name, "ms_per_episode": round(ms, 3), "episodes_per_sec": round(1e3 / ms, 2),
                      "query_frames_per_sec": round(qf * 1e3 / ms, 1)}), flush=True)


single = lambda m, c, cy: m.personalise(c, cy)
g = torch.Generator().manual_seed(1991)
# S1
m = prepare(orbit_b200.SingleStepFewShotRecogniser('resnet18', False, 'proto', 1, 256, False, 16), 84)
run("S1 ProtoNet+resnet18 84px 5-way 5-shot (25 support + 75 query frames)", m, [make_episode(EpisodeSpec(5, 5, 15, 1, 84), index=i) for i in range(4)], single, n=40)
del m
# S3
m = prepare(orbit_b200.SingleStepFewShotRecogniser('resnet18', True, 'versa', 1, 256, False, 16), 224)
ways = [int(w) for w in torch.randint(5, 16, (4,), generator=g)]
run(f"S3 CNAPs(versa+FiLM)+resnet18 224px ways={ways} 10-shot, 20 query clips/class", m,
    [make_episode(EpisodeSpec(w, 10, 20, 1, 224), index=10 + i) for i, w in enumerate(ways)], single)
del m
# S4
m = prepare(orbit_b200.MultiStepFewShotRecogniser('vit_b_32', False, 'linear', 1, 1024, False), 224)
def finetune(model, c, cy):
    model.personalise(c, cy, {'num_grad_steps': 50, 'learning_rate': 1e-3, 'optimizer': 'adam', 'loss_fn': None,
                              'extractor_lr_scale': 0.1, 'epsilon': 1e-8, 'weight_decay': 0.0, 'betas': (0.9, 0.999)})
run("S4 FineTuner+vit_b_32 224px 8-way 10-shot, 50 Adam steps, 160 query frames", m,
    [make_episode(EpisodeSpec(8, 10, 20, 1, 224), index=20 + i) for i in range(3)], finetune)
