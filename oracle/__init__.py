"""oracle/ -- CPU restatement of the ORBIT episodic hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package, and only as the *checker* or the
timed CPU baseline -- never as the product.  ``orbit_b200`` (the product) does not
import anything from here and fails loudly when its CUDA library is missing.

What is restated (plain PyTorch fp32 on CPU, written from scratch, each function citing
the reference file:line it follows; reference = microsoft/ORBIT-Dataset @ 97ccae1):

* ``backbones.py``  timm==0.6.12 ``tf_efficientnet_b0`` / ``vit_*_patch32_224`` (third-party
  dependency, ``requirements.txt:6``; its source is NOT under /root/reference -- the
  published architecture is restated with timm's state-dict key names) and the
  torchvision ``resnet18`` extension.
* ``parts.py``      pooler, prototype / linear / versa / mahalanobis heads, set encoder,
  FiLM generator (reference ``model/*.py``).
* ``recogniser.py`` the personalise()/predict() control flow (``model/few_shot_recognisers.py``).

Pinning status: the reference ships NO tests or golden vectors for this path
(SURVEY.md section 4), so the oracle is pinned against outputs of the *reference code
itself* run in the build container: ``oracle/make_golden.py`` imports the unmodified
reference modules from /root/reference (heads, pooler, set encoder, FiLM generator and the
whole ``few_shot_recognisers.py`` on top of the ``oracle/timm_shim`` stand-in for timm) and
writes ``tests/golden/*.npz``.  The backbone arithmetic itself (timm) is "parity unpinned"
against timm -- it is cross-checked structurally against torchvision's independent
``efficientnet_b0`` / ``vit_b_32`` / ``resnet18`` implementations instead (tests/test_oracle_backbones.py).
"""
