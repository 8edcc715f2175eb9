"""Put this directory on PYTHONPATH and the reference's scripts run on the aivc_b200 engine UNCHANGED:

    cd <AIVC>/src && PYTHONPATH=/path/to/aivc_b200_repo/compat python aivc.py -i ... --coding_config RA ...

`aivc.py` chains `encode.py` / `decode.py` / `evaluate.py` as SUBPROCESSES (src/aivc.py:117-139), so an in-process
monkey patch would be lost; Python imports `sitecustomize` at start-up of every interpreter that has this directory on
its path, which makes the shims subprocess-safe.  What it installs (aivc_b200/compat.py):

  * the layer mirrors under the reference's module paths (`layers.misc.custom_conv_layers`, ... -- a whole-module
    pickle `0_model.pt` stores module path + class name, model_management.py:347) and the `models` package the scrape
    lost (FullNet.GOP_forward = the fused CUDA encoder);
  * `torchac` (bitstream.py:10) backed by the C++ range coder of libaivc_b200.so;
  * `torch.set_deterministic` (cluster_mngt.py:37) and `torch.load(weights_only=False)` for whole-module pickles
    (model_management.py:347), both changed by torch >= 2.

Set AIVC_B200_NO_COMPAT=1 to switch it off for one process.
"""
import os
import sys

if not os.environ.get('AIVC_B200_NO_COMPAT'):
    _root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if _root not in sys.path:
        sys.path.insert(1, _root)
    try:
        from aivc_b200 import compat as _compat
        _compat.install()
    except Exception as _e:            # never break an unrelated interpreter start-up
        sys.stderr.write('aivc_b200 compat not installed: %r\n' % (_e,))
