from oracle.backbones import ConvBnAct, InvertedResidual, CondConvResidual, EdgeResidual  # noqa: F401
