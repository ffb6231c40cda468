"""``import speechmix`` -- the module name the reference's call sites use (ref:train.py:15 ``import speechmix``,
ref:eval.py, ref:eval.ipynb) -- resolves to the B200-native implementation: same class names
(``SpeechMixEED`` / ``SpeechMixFixed`` / ``SpeechMixAdapter`` / ``SpeechMixSelf`` and their ``HFSpeechMix*`` aliases,
``SpeechMixConfig``), same constructor and ``forward`` contract.  Everything lives in ``speechmix_b200``; this package
only re-exports it, so a maintainer switches by putting this repository on ``sys.path`` ahead of the reference
(INTEGRATION.md).  The s3prl / fairseq flavour (``ref:speechmix/model.py``) is out of scope and not provided."""
from speechmix_b200 import *  # noqa: F401,F403
from speechmix_b200 import __all__  # noqa: F401
