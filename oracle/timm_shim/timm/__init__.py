"""TEST-ONLY stand-in for the third-party ``timm`` package (timm==0.6.12 is pinned by the reference,
requirements.txt:6, but is not installed and cannot be installed offline).  It exposes just the
names the reference imports (model/feature_extractors.py:31-33, model/film.py:35-36,
utils/optim.py:6) and maps them onto the restatements in oracle/backbones.py, so that the
UNMODIFIED reference ``model/few_shot_recognisers.py`` can be imported from /root/reference by
oracle/make_golden.py.  Never imported by the product."""
__version__ = "0.6.12-shim"
